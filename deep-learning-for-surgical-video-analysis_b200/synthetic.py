"""Deterministic synthetic weights, inputs and video lengths for parity tests and the bench.

Workload definition shared by the bench, the smoke test and the parity tests.  Nothing here depends
on the reference or on the oracle, so the same tensors can be rebuilt on the GPU box (where
/root/reference is absent) and in this container (where the real reference consumes them to produce
tests/golden/*).

Weights are generated PER KEY from a generator seeded with (seed, crc32(key)), so they do not
depend on module construction order.  Two flavours:
  * "ref_init": the reference's own init distributions (mix_transformer_evp.py:300-313 — Linear
    trunc_normal(.02)/bias 0, LayerNorm 1/0, Conv2d N(0, sqrt(2/fan_out))/bias 0; torch defaults
    for flow_encoder / cross_attn / mstcn which the reference creates after `apply(_init_weights)`).
  * "stress": O(1) branch gains, non-zero biases, perturbed norm affines and BN statistics, so that
    an indexing / layout / bias bug anywhere changes the output by O(1).
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import numpy as np
import torch

# mit_bX_evp geometry (mix_transformer_evp.py:894-943); scale_factor 4 (:278)
EVP_CONFIGS = {
    "mit_b0_evp": dict(embed_dims=[32, 64, 160, 256], num_heads=[1, 2, 5, 8], depths=[2, 2, 2, 2], sr_ratios=[8, 4, 2, 1]),
    "mit_b1_evp": dict(embed_dims=[64, 128, 320, 512], num_heads=[1, 2, 5, 8], depths=[2, 2, 2, 2], sr_ratios=[8, 4, 2, 1]),
    "mit_b2_evp": dict(embed_dims=[64, 128, 320, 512], num_heads=[1, 2, 5, 8], depths=[3, 4, 6, 3], sr_ratios=[8, 4, 2, 1]),
    "mit_b3_evp": dict(embed_dims=[64, 128, 320, 512], num_heads=[1, 2, 5, 8], depths=[3, 4, 18, 3], sr_ratios=[8, 4, 2, 1]),
    "mit_b4_evp": dict(embed_dims=[64, 128, 320, 512], num_heads=[1, 2, 5, 8], depths=[3, 8, 27, 3], sr_ratios=[8, 4, 2, 1]),
    "mit_b5_evp": dict(embed_dims=[64, 128, 320, 512], num_heads=[1, 2, 5, 8], depths=[3, 6, 40, 3], sr_ratios=[8, 4, 2, 1]),
}
# dataset Normalize constants (generate_evp_LFB.py:247)
NORM_MEAN = (0.41757566, 0.26098573, 0.25888634)
NORM_STD = (0.21938758, 0.1983, 0.19342837)
# Cholec80 split sizes (generate_LFB_log.txt:8-10)
CHOLEC80_TRAIN_FRAMES = 86344
CHOLEC80_TEST_FRAMES = 98234


def evp_key_shapes(name: str = "mit_b3_evp", mlp_ratio: int = 4, embedding_dim: int = 2048) -> "OrderedDict[str, tuple]":
    """state_dict key -> shape for `mit_bX_evp()` (722 keys for b3)."""
    cfg = EVP_CONFIGS[name]
    dims, depths, srs = cfg["embed_dims"], cfg["depths"], cfg["sr_ratios"]
    out: "OrderedDict[str, tuple]" = OrderedDict()

    def lin(prefix, cin, cout, bias=True):
        out[prefix + ".weight"] = (cout, cin)
        if bias:
            out[prefix + ".bias"] = (cout,)

    def ln(prefix, c):
        out[prefix + ".weight"] = (c,)
        out[prefix + ".bias"] = (c,)

    def conv(prefix, cin, cout, k, groups=1, bias=True):
        out[prefix + ".weight"] = (cout, cin // groups, k, k)
        if bias:
            out[prefix + ".bias"] = (cout,)

    def bn(prefix, c):
        out[prefix + ".weight"] = (c,)
        out[prefix + ".bias"] = (c,)
        out[prefix + ".running_mean"] = (c,)
        out[prefix + ".running_var"] = (c,)
        out[prefix + ".num_batches_tracked"] = ()

    def patch(prefix, cin, cout, k):
        conv(prefix + ".proj", cin, cout, k)
        ln(prefix + ".norm", cout)

    ins = [3] + dims[:3]
    ks = [7, 3, 3, 3]
    for s in range(4):
        patch(f"patch_embed{s + 1}", ins[s], dims[s], ks[s])
    for s in range(4):
        c, h = dims[s], dims[s] * mlp_ratio
        for i in range(depths[s]):
            p = f"block{s + 1}.{i}"
            ln(p + ".norm1", c)
            lin(p + ".attn.q", c, c)
            lin(p + ".attn.kv", c, 2 * c)
            lin(p + ".attn.proj", c, c)
            if srs[s] > 1:
                conv(p + ".attn.sr", c, c, srs[s])
                ln(p + ".attn.norm", c)
            ln(p + ".norm2", c)
            lin(p + ".mlp.fc1", c, h)
            conv(p + ".mlp.dwconv.dwconv", h, h, 3, groups=h)
            lin(p + ".mlp.fc2", h, c)
        ln(f"norm{s + 1}", c)
    # head (segformer_head.py:66-106); creation order c4,c3,c2,c1
    for i in (4, 3, 2, 1):
        lin(f"head.linear_c{i}.proj", dims[i - 1], embedding_dim)
    out["head.linear_fuse.conv.weight"] = (embedding_dim, 4 * embedding_dim, 1, 1)
    bn("head.linear_fuse.bn", embedding_dim)
    for nm in ("fc", "fc_ant"):
        lin(f"head.{nm}.0", 2048, 512)
        lin(f"head.{nm}.2", 512, 7)
    # prompt generator (mix_transformer_evp.py:580-642)
    pins = [3] + [d // 4 for d in dims[:3]]
    for s in range(4):
        patch(f"prompt_generator.handcrafted_generator{s + 1}", pins[s], dims[s] // 4, ks[s])
    for s in range(4):
        lin(f"prompt_generator.embedding_generator{s + 1}", dims[s], dims[s] // 4)
    for s in range(4):
        for i in range(depths[s]):
            lin(f"prompt_generator.lightweight_mlp{s + 1}_{i}.0", dims[s] // 4, dims[s] // 4)
        lin(f"prompt_generator.shared_mlp{s + 1}", dims[s] // 4, dims[s])
    # flow encoder (:823-836)
    fch = [2, 64, 128, dims[2], dims[3]]
    for i in range(4):
        conv(f"flow_encoder.conv{i + 1}", fch[i], fch[i + 1], 7 if i == 0 else 3)
        bn(f"flow_encoder.bn{i + 1}", fch[i + 1])
    # cross attention (:868-876)
    for s in (3, 4):
        c = dims[s - 1]
        out[f"cross_attn_s{s}.cross_attn.in_proj_weight"] = (3 * c, c)
        out[f"cross_attn_s{s}.cross_attn.in_proj_bias"] = (3 * c,)
        lin(f"cross_attn_s{s}.cross_attn.out_proj", c, c)
        ln(f"cross_attn_s{s}.norm", c)
    return out


def mstcn_key_shapes(stages=2, layers=8, f_maps=32, f_dim=2048, out_features=14) -> "OrderedDict[str, tuple]":
    """state_dict key -> shape for `MultiStageModel_S` (mstcn.py:94-120, 153-171, 181-206); 72 keys."""
    out: "OrderedDict[str, tuple]" = OrderedDict()

    def stage(prefix, dim):
        out[prefix + ".conv_1x1.weight"] = (f_maps, dim, 1)
        out[prefix + ".conv_1x1.bias"] = (f_maps,)
        for i in range(layers):
            out[f"{prefix}.layers.{i}.conv_dilated.weight"] = (f_maps, f_maps, 3)
            out[f"{prefix}.layers.{i}.conv_dilated.bias"] = (f_maps,)
            out[f"{prefix}.layers.{i}.conv_1x1.weight"] = (f_maps, f_maps, 1)
            out[f"{prefix}.layers.{i}.conv_1x1.bias"] = (f_maps,)
        out[prefix + ".conv_out_classes.weight"] = (out_features, f_maps, 1)
        out[prefix + ".conv_out_classes.bias"] = (out_features,)

    stage("stage1_phase", f_dim)
    for s in range(stages - 1):
        stage(f"stages.{s}", out_features)
    return out


def _gen(seed: int, key: str) -> torch.Generator:
    return torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2**63 - 1))


def _is_norm(key: str) -> bool:
    parts = key.split(".")
    owner = parts[-2] if len(parts) >= 2 else ""
    return owner.startswith("norm") or owner.startswith("bn")


def synth_state_dict(shapes: "OrderedDict[str, tuple]", seed: int = 0, mode: str = "ref_init") -> "OrderedDict[str, torch.Tensor]":
    assert mode in ("ref_init", "stress")
    stress = mode == "stress"
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, shape in shapes.items():
        g = _gen(seed, key)
        leaf = key.split(".")[-1]
        if leaf == "num_batches_tracked":
            t = torch.zeros((), dtype=torch.long)
        elif leaf == "running_mean":
            t = 0.2 * torch.randn(shape, generator=g) if stress else torch.zeros(shape)
        elif leaf == "running_var":
            t = 0.5 + torch.rand(shape, generator=g) if stress else torch.ones(shape)
        elif _is_norm(key):
            if leaf == "weight":
                t = 1.0 + (0.1 * torch.randn(shape, generator=g) if stress else 0.0) * torch.ones(shape)
            else:
                t = 0.1 * torch.randn(shape, generator=g) if stress else torch.zeros(shape)
        elif leaf in ("weight", "in_proj_weight"):
            torch_default = key.startswith(("flow_encoder", "cross_attn", "stage1_phase", "stages.")) or key.startswith("head.fc")
            if len(shape) == 2:  # Linear [out, in]
                fan_in = shape[1]
                if stress:
                    t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
                elif torch_default:
                    b = 1.0 / math.sqrt(fan_in)
                    t = (torch.rand(shape, generator=g) * 2 - 1) * b
                else:
                    t = torch.randn(shape, generator=g).clamp_(-100.0, 100.0) * 0.02  # trunc at +-2 abs never binds at std .02
            elif len(shape) == 4:  # Conv2d [out, in/groups, kh, kw]
                fan_in = shape[1] * shape[2] * shape[3]
                fan_out = shape[0] * shape[2] * shape[3]
                depthwise = shape[1] == 1 and shape[0] > 1 and "dwconv" in key
                if depthwise:
                    fan_out = shape[2] * shape[3]
                if stress:
                    t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
                elif torch_default:
                    b = 1.0 / math.sqrt(fan_in)
                    t = (torch.rand(shape, generator=g) * 2 - 1) * b
                else:
                    t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_out)
            elif len(shape) == 3:  # Conv1d [out, in, k]
                fan_in = shape[1] * shape[2]
                b = (2.0 if stress else 1.0) / math.sqrt(fan_in)
                t = (torch.rand(shape, generator=g) * 2 - 1) * b
            else:
                raise ValueError(key)
        elif leaf in ("bias", "in_proj_bias"):
            if stress:
                t = 0.1 * torch.randn(shape, generator=g)
            elif key.startswith(("flow_encoder", "stage1_phase", "stages.")) or key.startswith("head.fc"):
                t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
            else:
                t = torch.zeros(shape)
        else:
            raise ValueError(key)
        sd[key] = t.to(torch.float32) if t.dtype != torch.long else t
    return sd


def synth_frames(n: int, seed: int, H: int = 224, W: int = 224, device="cpu"):
    """Synthetic (frames, segmaps, flow) of the shapes the LFB driver feeds the model
    (generate_evp_LFB.py:448-451): x ~ N(0,1); seg = Normalize(binary mask, p=0.3, same in 3 channels)
    (data_process.py:417); flow ~ 2*N(0,1) pixels."""
    g = torch.Generator().manual_seed(1_000_000 + seed)
    x = torch.randn(n, 1, 3, H, W, generator=g)
    mask = (torch.rand(n, 1, 1, H, W, generator=g) > 0.7).float().expand(n, 1, 3, H, W)
    mean = torch.tensor(NORM_MEAN).view(1, 1, 3, 1, 1)
    std = torch.tensor(NORM_STD).view(1, 1, 3, 1, 1)
    seg = ((mask - mean) / std).contiguous()
    flow = 2.0 * torch.randn(n, 1, 2, H, W, generator=g)
    return x.to(device), seg.to(device), flow.to(device)


def cholec80_video_lengths() -> np.ndarray:
    """80 deterministic video lengths: lognormal(7.6, .45) clipped to [700, 6000], rescaled so that videos
    0-39 sum to 86 344 frames and 40-79 to 98 234."""
    rng = np.random.default_rng(80)
    raw = np.clip(rng.lognormal(7.6, 0.45, size=80), 700, 6000)
    out = np.zeros(80, dtype=np.int64)
    for lo, hi, total in ((0, 40, CHOLEC80_TRAIN_FRAMES), (40, 80, CHOLEC80_TEST_FRAMES)):
        part = raw[lo:hi] * (total / raw[lo:hi].sum())
        ints = np.floor(part).astype(np.int64)
        rem = total - ints.sum()
        order = np.argsort(-(part - ints))
        ints[order[:rem]] += 1
        out[lo:hi] = ints
    return out


def _phase_protos(dim: int = 2048) -> torch.Tensor:
    return torch.randn(7, dim, generator=torch.Generator().manual_seed(424242))


def synth_lfb_features(T: int, seed: int, dim: int = 2048) -> torch.Tensor:
    """Synthetic LFB rows [T, dim] fp32: positive, mean ~0.285 like the real pooled-ReLU features
    (BASELINE.md §2), piecewise-stationary over 7 pseudo-phases (prototypes shared by all videos) so that
    the per-frame argmax of the MS-TCN logits is not a constant class."""
    g = torch.Generator().manual_seed(2_000_000 + seed)
    protos = _phase_protos(dim)
    order = torch.randperm(7, generator=g)
    bounds = torch.sort(torch.rand(6, generator=g)).values
    t = (torch.arange(T, dtype=torch.float32) + 0.5) / max(T, 1)
    phase = order[(t[:, None] > bounds[None, :]).sum(dim=1)]
    noise = torch.randn(T, dim, generator=g)
    return (0.285 * (1.0 + 0.6 * protos[phase] + 0.4 * noise)).abs().contiguous()


def synth_mstcn_state_dict(stages=2, layers=8, f_maps=32, f_dim=2048, out_features=14, seed: int = 1, mode: str = "ref_init"):
    """MS-TCN weights. mode "phase" = "stress" weights at half gain plus a structured component that makes
    class c (<7) respond to pseudo-phase c of `synth_lfb_features`, giving a non-degenerate argmax histogram
    with realistic top-2 margins for the argmax-agreement parity gate."""
    shapes = mstcn_key_shapes(stages, layers, f_maps, f_dim, out_features)
    if mode != "phase":
        return synth_state_dict(shapes, seed, mode)
    sd = synth_state_dict(shapes, seed, "stress")
    for k in sd:
        if k.endswith("weight"):
            sd[k] = sd[k] * 0.5
    protos = _phase_protos(f_dim)
    w = sd["stage1_phase.conv_1x1.weight"]
    w[:7, :, 0] += 3.0 * protos / (0.285 * 0.6 * f_dim)
    for p, gain_in in (("stage1_phase", None), ) + tuple((f"stages.{s}", 3.0) for s in range(stages - 1)):
        if gain_in is not None:
            for c in range(7):
                sd[p + ".conv_1x1.weight"][c, c, 0] += gain_in
        for c in range(7):
            sd[p + ".conv_out_classes.weight"][c, c, 0] += 1.5
    return sd


# ----------------------------------------------------------------------------------------------- chained (encoder -> MS-TCN) gate
def phase_schedule(T: int, seed: int) -> torch.Tensor:
    """Piecewise-constant pseudo-phase id (0..6) per frame: a random order of the 7 phases with random boundaries, as in
    `synth_lfb_features`."""
    g = torch.Generator().manual_seed(3_000_000 + seed)
    order = torch.randperm(7, generator=g)
    bounds = torch.sort(torch.rand(6, generator=g)).values
    t = (torch.arange(T, dtype=torch.float32) + 0.5) / max(T, 1)
    return order[(t[:, None] > bounds[None, :]).sum(dim=1)]


def synth_phase_frames(T: int, seed: int, H: int = 224, W: int = 224, amp: float = 1.0, phase: "torch.Tensor | None" = None):
    """`synth_frames` plus a per-phase low-frequency image pattern added to the frames (7 fixed 7x7 random fields, bilinearly
    upsampled), so that the ENCODER's features carry a phase signal and the chained LFB -> MS-TCN argmax is not a constant class.
    Returns (x, seg, flow, phase)."""
    x, seg, flow = synth_frames(T, seed, H, W)
    if phase is None:
        phase = phase_schedule(T, seed)
    pat = torch.nn.functional.interpolate(torch.randn(7, 3, 7, 7, generator=torch.Generator().manual_seed(99)), size=(H, W), mode="bilinear",
                                          align_corners=False)
    return x + amp * pat[phase].unsqueeze(1), seg, flow, phase


def synth_mstcn_state_dict_for_protos(dev_protos: torch.Tensor, center: torch.Tensor, gain: float = 3.0, stages=2, layers=8, f_maps=32, out_features=14,
                                      seed: int = 1):
    """MS-TCN "phase" weights whose class c (< 7) responds to feature deviation `dev_protos[c]` from `center` (both measured on a
    calibration set of encoder features): stage-1 projection row c += gain * d_c / |d_c|^2, bias compensates the centre."""
    f_dim = dev_protos.shape[1]
    sd = synth_state_dict(mstcn_key_shapes(stages, layers, f_maps, f_dim, out_features), seed, "stress")
    for k in sd:
        if k.endswith("weight"):
            sd[k] = sd[k] * 0.5
    w = gain * dev_protos / (dev_protos.norm(dim=1, keepdim=True) ** 2)
    sd["stage1_phase.conv_1x1.weight"][:7, :, 0] += w
    sd["stage1_phase.conv_1x1.bias"][:7] -= w @ center
    for p, gain_in in (("stage1_phase", None), ) + tuple((f"stages.{s}", 3.0) for s in range(stages - 1)):
        if gain_in is not None:
            for c in range(7):
                sd[p + ".conv_1x1.weight"][c, c, 0] += gain_in
        for c in range(7):
            sd[p + ".conv_out_classes.weight"][c, c, 0] += 1.5
    return sd

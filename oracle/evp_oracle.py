"""CPU/fp32 restatement of the reference LFB path: `mit_bX_evp.forward(x, seg, flow, return_features=True)`.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this file; the product package never does.

It is a *functional* restatement over a plain `state_dict` (no nn.Module tree), in plain PyTorch fp32,
each function citing the reference lines it follows (paths relative to /root/reference).  Pinning:
`tests/test_oracle_cpu.py` checks it (a) against the real reference imported through
`oracle/ref_loader.py` when /root/reference exists, and (b) everywhere against the committed golden
vectors in `tests/golden/` which were produced by the real reference (`tests/golden/make_golden.py`).
The reference has no tests or golden vectors of its own, so outputs of the reference
itself, run in the build container, are the anchor.

`emulate_bf16=True` rounds every GEMM/conv operand (activations and weights) to bf16 before the
contraction, keeping fp32 accumulation and an fp32 residual stream — the arithmetic the CUDA path
uses.  It is used to calibrate tolerances, never as the expected value.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


class _Ctx:
    def __init__(self, emulate_bf16: bool):
        self.bf16 = emulate_bf16

    def r(self, t: torch.Tensor) -> torch.Tensor:
        return t.to(torch.bfloat16).to(torch.float32) if self.bf16 else t


def _linear(cx: _Ctx, sd: SD, p: str, x: torch.Tensor, bias: bool = True) -> torch.Tensor:
    return F.linear(cx.r(x), cx.r(sd[p + ".weight"]), sd[p + ".bias"] if bias else None)


def _ln(sd: SD, p: str, x: torch.Tensor, eps: float) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def overlap_patch_embed(cx, sd, p, x_nchw, k, stride):
    """OverlapPatchEmbed.forward (mix_transformer_evp.py:209-215): conv(k, stride, pad k//2) -> tokens -> LN(1e-5)."""
    y = F.conv2d(cx.r(x_nchw), cx.r(sd[p + ".proj.weight"]), sd[p + ".proj.bias"], stride=stride, padding=k // 2)
    _, _, H, W = y.shape
    y = y.flatten(2).transpose(1, 2)
    return _ln(sd, p + ".norm", y, 1e-5), H, W


def attention(cx, sd, p, x, H, W, heads, sr):
    """Attention.forward (mix_transformer_evp.py:110-131)."""
    B, N, C = x.shape
    hd = C // heads
    q = _linear(cx, sd, p + ".q", x).reshape(B, N, heads, hd).permute(0, 2, 1, 3)
    if sr > 1:
        x_ = x.permute(0, 2, 1).reshape(B, C, H, W)
        x_ = F.conv2d(cx.r(x_), cx.r(sd[p + ".sr.weight"]), sd[p + ".sr.bias"], stride=sr)
        x_ = x_.reshape(B, C, -1).permute(0, 2, 1)
        x_ = _ln(sd, p + ".norm", x_, 1e-5)
    else:
        x_ = x
    kv = _linear(cx, sd, p + ".kv", x_).reshape(B, -1, 2, heads, hd).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    attn = (cx.r(q) @ cx.r(k).transpose(-2, -1)) * (hd ** -0.5)
    attn = attn.softmax(dim=-1)
    o = (cx.r(attn) @ cx.r(v)).transpose(1, 2).reshape(B, N, C)
    return _linear(cx, sd, p + ".proj", o)


def mix_ffn(cx, sd, p, x, H, W):
    """Mlp.forward + DWConv.forward (mix_transformer_evp.py:60-67, 24-30): fc1 -> dw3x3 -> GELU(erf) -> fc2."""
    B, N, _ = x.shape
    h = _linear(cx, sd, p + ".fc1", x)
    Ch = h.shape[-1]
    hh = cx.r(h).transpose(1, 2).reshape(B, Ch, H, W)
    hh = F.conv2d(hh, sd[p + ".dwconv.dwconv.weight"], sd[p + ".dwconv.dwconv.bias"], stride=1, padding=1, groups=Ch)
    h = F.gelu(hh.flatten(2).transpose(1, 2))
    return _linear(cx, sd, p + ".fc2", h)


def block(cx, sd, p, x, H, W, heads, sr):
    """Block.forward (mix_transformer_evp.py:167-171); norm eps 1e-6 (:898-926)."""
    x = x + attention(cx, sd, p + ".attn", _ln(sd, p + ".norm1", x, 1e-6), H, W, heads, sr)
    x = x + mix_ffn(cx, sd, p + ".mlp", _ln(sd, p + ".norm2", x, 1e-6), H, W)
    return x


def gaussian_filter(seg_nchw):
    """GaussianFilter.conv_gauss (mix_transformer_evp.py:500-514): reflect pad 2 + 5x5 binomial /256, depthwise."""
    k1 = torch.tensor([1.0, 4.0, 6.0, 4.0, 1.0], dtype=seg_nchw.dtype, device=seg_nchw.device)
    k = (k1[:, None] * k1[None, :]) / 256.0
    C = seg_nchw.shape[1]
    img = F.pad(seg_nchw, (2, 2, 2, 2), mode="reflect")
    return F.conv2d(img, k.expand(C, 1, 5, 5).contiguous(), groups=C)


def handcrafted_prompts(cx, sd, seg_nchw):
    """PromptGenerator.init_prompts (mix_transformer_evp.py:718-747)."""
    x = gaussian_filter(seg_nchw)
    B = x.shape[0]
    feats = []
    ks, st = [7, 3, 3, 3], [4, 2, 2, 2]
    for s in range(4):
        t, H, W = overlap_patch_embed(cx, sd, f"prompt_generator.handcrafted_generator{s + 1}", x, ks[s], st[s])
        feats.append(t)
        x = t.reshape(B, H, W, -1).permute(0, 3, 1, 2).contiguous()
    return feats


def adapter(cx, sd, x, prompt_sum, s, i):
    """PromptGenerator.get_prompt, adaptor branch (mix_transformer_evp.py:776-815)."""
    f = F.gelu(_linear(cx, sd, f"prompt_generator.lightweight_mlp{s}_{i}.0", prompt_sum))
    return x + _linear(cx, sd, f"prompt_generator.shared_mlp{s}", f)


def forward_features(cx, sd, cfg, x_nchw, seg_nchw, taps=None):
    """MixVisionTransformerEVP.forward_features (mix_transformer_evp.py:352-416) for any HxW (F8)."""
    dims, heads, depths, srs = cfg["embed_dims"], cfg["num_heads"], cfg["depths"], cfg["sr_ratios"]
    B = x_nchw.shape[0]
    hc = handcrafted_prompts(cx, sd, seg_nchw)
    outs = []
    ks, st = [7, 3, 3, 3], [4, 2, 2, 2]
    x = x_nchw
    for s in range(4):
        x, H, W = overlap_patch_embed(cx, sd, f"patch_embed{s + 1}", x, ks[s], st[s])
        # init_prompt (:749-756): embedding_generator on the patch-embed output, constant across depth
        psum = hc[s] + _linear(cx, sd, f"prompt_generator.embedding_generator{s + 1}", x)
        for i in range(depths[s]):
            x = adapter(cx, sd, x, psum, s + 1, i)
            x = block(cx, sd, f"block{s + 1}.{i}", x, H, W, heads[s], srs[s])
        x = _ln(sd, f"norm{s + 1}", x, 1e-6)
        if taps is not None:
            taps[f"stage{s + 1}_tokens"] = x  # [B, N, C] == NHWC
        x = x.reshape(B, H, W, -1).permute(0, 3, 1, 2).contiguous()
        outs.append(x)
    return outs


def _bn_eval(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, 1e-5)


def flow_encoder(cx, sd, flow_nchw):
    """OpticalFlowEncoder.forward (mix_transformer_evp.py:838-859)."""
    x = flow_nchw
    strides, pads = [4, 2, 2, 2], [3, 1, 1, 1]
    feats = []
    for i in range(4):
        x = F.conv2d(cx.r(x), cx.r(sd[f"flow_encoder.conv{i + 1}.weight"]), sd[f"flow_encoder.conv{i + 1}.bias"], stride=strides[i], padding=pads[i])
        x = F.relu(_bn_eval(sd, f"flow_encoder.bn{i + 1}", x))
        feats.append(x)
    return feats[2].flatten(2).transpose(1, 2), feats[3].flatten(2).transpose(1, 2)


def cross_attention(cx, sd, p, xv, xf, heads=8):
    """MotionGuidedCrossAttention.forward (mix_transformer_evp.py:878-890) with nn.MultiheadAttention
    semantics: in_proj_weight = [Wq;Wk;Wv], scale hd^-0.5, out_proj, LN(x_visual + attn) eps 1e-5."""
    B, N, C = xv.shape
    hd = C // heads
    W, b = sd[p + ".cross_attn.in_proj_weight"], sd[p + ".cross_attn.in_proj_bias"]
    q = F.linear(cx.r(xv), cx.r(W[:C]), b[:C]).reshape(B, N, heads, hd).transpose(1, 2)
    k = F.linear(cx.r(xf), cx.r(W[C:2 * C]), b[C:2 * C]).reshape(B, -1, heads, hd).transpose(1, 2)
    v = F.linear(cx.r(xf), cx.r(W[2 * C:]), b[2 * C:]).reshape(B, -1, heads, hd).transpose(1, 2)
    a = ((cx.r(q) @ cx.r(k).transpose(-2, -1)) * (hd ** -0.5)).softmax(-1)
    o = (cx.r(a) @ cx.r(v)).transpose(1, 2).reshape(B, N, C)
    o = _linear(cx, sd, p + ".cross_attn.out_proj", o)
    return _ln(sd, p + ".norm", xv + o, 1e-5)


def segformer_head_features(cx, sd, outs):
    """SegFormerHead.forward(..., return_features=True) (segformer_head.py:137-173)."""
    c1, c2, c3, c4 = outs
    n, _, h, w = c4.shape
    cat = []
    for i, c in ((4, c4), (3, c3), (2, c2), (1, c1)):
        t = _linear(cx, sd, f"head.linear_c{i}.proj", c.flatten(2).transpose(1, 2))
        t = t.permute(0, 2, 1).reshape(n, -1, c.shape[2], c.shape[3])
        if i != 4:
            t = F.interpolate(t, size=(h, w), mode="bilinear", align_corners=False)
        cat.append(t)
    x = F.conv2d(cx.r(torch.cat(cat, dim=1)), cx.r(sd["head.linear_fuse.conv.weight"]))
    x = F.relu(_bn_eval(sd, "head.linear_fuse.bn", x))
    return x.mean(dim=(2, 3))  # Dropout2d = id in eval; AdaptiveAvgPool2d(1); flatten


def head_logits(sd, feats):
    """head.fc / head.fc_ant (segformer_head.py:101-106, 176-179)."""
    def mlp(p):
        return F.linear(F.relu(F.linear(feats, sd[p + ".0.weight"], sd[p + ".0.bias"])), sd[p + ".2.weight"], sd[p + ".2.bias"])
    return mlp("head.fc"), mlp("head.fc_ant")


@torch.no_grad()
def evp_forward(sd: SD, cfg: dict, x: torch.Tensor, seg: torch.Tensor, flow: Optional[torch.Tensor] = None,
                return_features: bool = True, emulate_bf16: bool = False, taps: Optional[dict] = None):
    """MixVisionTransformerEVP.forward (mix_transformer_evp.py:418-449).
    x, seg: [B,1,3,H,W] (or [B,3,H,W]); flow: [B,1,2,H,W] or None."""
    cx = _Ctx(emulate_bf16)
    H, W = x.shape[-2], x.shape[-1]
    x = x.reshape(-1, 3, H, W).float()
    seg = seg.reshape(-1, 3, H, W).float()
    outs = forward_features(cx, sd, cfg, x, seg, taps)
    if flow is not None:
        f3, f4 = flow_encoder(cx, sd, flow.reshape(-1, 2, H, W).float())
        for idx, p, ft in ((2, "cross_attn_s3", f3), (3, "cross_attn_s4", f4)):
            c = outs[idx]
            Bc, Cc, Hc, Wc = c.shape
            fused = cross_attention(cx, sd, p, c.flatten(2).transpose(1, 2), ft)
            if taps is not None:
                taps[f"fused{idx + 1}_tokens"] = fused
            outs[idx] = fused.transpose(1, 2).reshape(Bc, Cc, Hc, Wc)
    feats = segformer_head_features(cx, sd, outs)
    if return_features:
        return feats
    return head_logits(sd, feats)

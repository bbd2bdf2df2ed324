"""CPU/fp32 restatement of `mstcn.MultiStageModel_S.forward` (reference: /root/reference/mstcn.py).

TEST INFRASTRUCTURE ONLY (same rules as evp_oracle.py).  Functional over a plain state_dict.
Pinned against the real reference (when present) and against tests/golden/mstcn_*.npz, which were
produced by the real reference; additionally against the survey's known answer
(`MultiStageModel_S(2,8,32,2048,14,True)` at torch.manual_seed(0): sum = -4348.2007).

Two equivalent forms are given: `mstcn_forward` (torch conv1d, literal) and `mstcn_forward_taps`
(explicit causal-tap formulation y[t] = sum_k W_k x[t-(2-k)d], the form the CUDA kernel implements).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _n_layers(sd: SD, prefix: str) -> int:
    n = 0
    while f"{prefix}.layers.{n}.conv_dilated.weight" in sd:
        n += 1
    return n


def dilated_residual_layer(sd: SD, p: str, x: torch.Tensor, d: int) -> torch.Tensor:
    """DilatedResidualLayer.forward, causal branch (mstcn.py:208-214): conv(k3, pad 2d, dil d) -> relu ->
    drop last 2d samples -> conv 1x1 -> dropout (eval: id) -> x + out."""
    out = F.relu(F.conv1d(x, sd[p + ".conv_dilated.weight"], sd[p + ".conv_dilated.bias"], padding=2 * d, dilation=d))
    out = out[:, :, : -(2 * d)]
    out = F.conv1d(out, sd[p + ".conv_1x1.weight"], sd[p + ".conv_1x1.bias"])
    return x + out


def single_stage(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """SingleStageModel.forward (mstcn.py:173-178)."""
    out = F.conv1d(x, sd[p + ".conv_1x1.weight"], sd[p + ".conv_1x1.bias"])
    for i in range(_n_layers(sd, p)):
        out = dilated_residual_layer(sd, f"{p}.layers.{i}", out, 2 ** i)
    return F.conv1d(out, sd[p + ".conv_out_classes.weight"], sd[p + ".conv_out_classes.bias"])


@torch.no_grad()
def mstcn_forward(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """MultiStageModel_S.forward (mstcn.py:122-130). x: [1, dim, T] -> [stages, 1, C, T]."""
    out = single_stage(sd, "stage1_phase", x.float())
    outs = [out]
    s = 0
    while f"stages.{s}.conv_1x1.weight" in sd:
        out = single_stage(sd, f"stages.{s}", F.softmax(out, dim=1))
        outs.append(out)
        s += 1
    return torch.stack(outs, dim=0)


@torch.no_grad()
def mstcn_forward_taps(sd: SD, feats_tm: torch.Tensor) -> torch.Tensor:
    """Same function in the time-major / causal-tap form. feats_tm: [T, dim] -> [stages, C, T]."""
    def stage(p, x):  # x: [T, dim]
        h = x @ sd[p + ".conv_1x1.weight"][:, :, 0].t() + sd[p + ".conv_1x1.bias"]
        T = h.shape[0]
        for i in range(_n_layers(sd, p)):
            d = 2 ** i
            w = sd[f"{p}.layers.{i}.conv_dilated.weight"]  # [out, in, 3]
            acc = sd[f"{p}.layers.{i}.conv_dilated.bias"].expand(T, -1).clone()
            for k in range(3):
                sh = (2 - k) * d
                if sh >= T:
                    continue
                xs = torch.zeros_like(h)
                xs[sh:] = h[: T - sh] if sh > 0 else h
                acc = acc + xs @ w[:, :, k].t()
            acc = F.relu(acc)
            h = h + acc @ sd[f"{p}.layers.{i}.conv_1x1.weight"][:, :, 0].t() + sd[f"{p}.layers.{i}.conv_1x1.bias"]
        return h @ sd[p + ".conv_out_classes.weight"][:, :, 0].t() + sd[p + ".conv_out_classes.bias"]

    out = stage("stage1_phase", feats_tm.float())
    outs = [out]
    s = 0
    while f"stages.{s}.conv_1x1.weight" in sd:
        out = stage(f"stages.{s}", F.softmax(out, dim=1))
        outs.append(out)
        s += 1
    return torch.stack([o.t() for o in outs], dim=0)

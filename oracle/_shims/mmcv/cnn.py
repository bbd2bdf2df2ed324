"""Shim for `mmcv.cnn` (reference import: segformer_head.py:4).

`ConvModule(in, out, k, norm_cfg=dict(type='BN'))` as called at segformer_head.py:74-80 means
Conv2d(bias=False) -> BatchNorm2d (submodule `bn`) -> ReLU(inplace) (submodule `activate`),
order conv/norm/act; state_dict keys `conv.weight`, `bn.*`.
"""
import torch.nn as nn


class ConvModule(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias="auto", conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU"),
                 inplace=True, **kwargs):
        super().__init__()
        with_norm = norm_cfg is not None
        if bias == "auto":
            bias = not with_norm
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                              dilation=dilation, groups=groups, bias=bias)
        self.with_norm = with_norm
        if with_norm:
            assert norm_cfg.get("type") in ("BN", "SyncBN")
            self.bn = nn.BatchNorm2d(out_channels)
        self.with_activation = act_cfg is not None
        if self.with_activation:
            assert act_cfg.get("type") == "ReLU"
            self.activate = nn.ReLU(inplace=inplace)

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = self.bn(x)
        if self.with_activation:
            x = self.activate(x)
        return x


class DepthwiseSeparableConvModule(nn.Module):  # imported by the reference, never instantiated
    def __init__(self, *a, **k):
        raise NotImplementedError("not used on the LFB path")

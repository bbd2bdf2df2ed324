"""Drop-in parameter holder for the reference's `models/segformer_head.py` (SegFormerHead, MLP).

Same submodule names / state_dict keys (`linear_c{1..4}.proj.*`, `linear_fuse.{conv,bn}.*`, `fc.{0,2}.*`, `fc_ant.{0,2}.*`;
segformer_head.py:66-106).  `mmcv.cnn.ConvModule(norm_cfg=BN)` is restated as conv(bias=False) + bn + activate.
The arithmetic (segformer_head.py:137-179) runs in libsurgvid.so.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from .. import _native


class _Holder(nn.Module):
    """A module that only owns parameters; its math lives in the fused CUDA forward."""

    def forward(self, *args, **kwargs):
        raise RuntimeError(f"{type(self).__name__} is a parameter holder: its computation is fused into the surgvid CUDA "
                           "forward (no PyTorch fallback exists)")


class MLP(_Holder):
    def __init__(self, input_dim=2048, embed_dim=768):
        super().__init__()
        self.proj = nn.Linear(input_dim, embed_dim)


class _ConvBNReLU(_Holder):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1, bias=False)
        self.bn = nn.BatchNorm2d(out_channels)
        self.activate = nn.ReLU(inplace=True)


class SegFormerHead(_Holder):
    def __init__(self, in_channels, num_classes):
        super().__init__()
        self.input_transform, self.in_index, self.align_corners = "multiple_select", [0, 1, 2, 3], False
        self.embedding_dim = self.embedding_dim1 = 2048
        self.in_channels, self.num_classes = in_channels, num_classes
        self.dropout = nn.Dropout2d(0.1)
        c1, c2, c3, c4 = in_channels
        self.linear_c4 = MLP(input_dim=c4, embed_dim=self.embedding_dim)
        self.linear_c3 = MLP(input_dim=c3, embed_dim=self.embedding_dim)
        self.linear_c2 = MLP(input_dim=c2, embed_dim=self.embedding_dim)
        self.linear_c1 = MLP(input_dim=c1, embed_dim=self.embedding_dim)
        self.linear_fuse = _ConvBNReLU(self.embedding_dim * 4, self.embedding_dim)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Sequential(nn.Linear(2048, 512), nn.ReLU(), nn.Linear(512, 7))
        self.fc_ant = nn.Sequential(nn.Linear(2048, 512), nn.ReLU(), nn.Linear(512, 7))

    def classify(self, model, feats: torch.Tensor):
        """(y, y_ant) = (fc(feats), fc_ant(feats)) — segformer_head.py:176-179 — on the device, fp32."""
        st = model._state(feats.device)
        B = feats.shape[0]
        y = torch.empty((B, 7), dtype=torch.float32, device=feats.device)
        y_ant = torch.empty_like(y)
        feats = feats.contiguous()
        rc = _native.lib().sv_evp_classify(st["handle"], ctypes.c_void_p(feats.data_ptr()), ctypes.c_void_p(y.data_ptr()),
                                           ctypes.c_void_p(y_ant.data_ptr()), B, ctypes.c_void_p(torch.cuda.current_stream(feats.device).cuda_stream))
        _native.check(rc, "sv_evp_classify")
        return y, y_ant

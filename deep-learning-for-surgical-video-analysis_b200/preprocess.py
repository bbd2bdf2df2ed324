"""Device-side replacement of the reference's per-frame input transforms.

The reference decodes a JPEG, then on the CPU inside `CholecFlowDataset.__getitem__` (data_process.py:409-483) applies
`Resize((250,250)) -> CenterCrop(224) -> ToTensor -> Normalize` to the frame and to its segmentation map
(generate_evp_LFB.py:242-248) and `cv2.resize` + displacement rescale + crop to the RAFT flow.  Here the raw uint8 frames
and the raw float32 flow are copied to the GPU and the same arithmetic runs there (csrc/preprocess.cu): images come out
bit-identical to torchvision's, flow to within float32 rounding of OpenCV's.  CUDA only, no CPU implementation.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _native

MEAN = (0.41757566, 0.26098573, 0.25888634)  # generate_evp_LFB.py:247
STD = (0.21938758, 0.1983, 0.19342837)


class FramePreprocessor:
    """frames/segmaps: uint8 [B, H, W, 3] (RGB, HWC — the layout a decoded frame has); flow: float32 [B, Hf, Wf, 2]."""

    def __init__(self, in_hw: Tuple[int, int], flow_hw: Optional[Tuple[int, int]] = None, resize: int = 250, crop: int = 224,
                 mean: Sequence[float] = MEAN, std: Sequence[float] = STD):
        self.in_hw = (int(in_hw[0]), int(in_hw[1]))
        self.flow_hw = None if flow_hw is None else (int(flow_hw[0]), int(flow_hw[1]))
        self.resize, self.crop = int(resize), int(crop)
        lib = _native.lib()
        m = (ctypes.c_float * 3)(*[float(v) for v in mean])
        s = (ctypes.c_float * 3)(*[float(v) for v in std])
        h = ctypes.c_void_p()
        fh, fw = self.flow_hw if self.flow_hw is not None else (0, 0)
        _native.check(lib.sv_prep_create(self.in_hw[0], self.in_hw[1], fh, fw, self.resize, self.crop, m, s, ctypes.byref(h)), "sv_prep_create")
        self._h = h
        self._ws: Optional[torch.Tensor] = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _native.lib().sv_prep_destroy(h)
            except Exception:
                pass
            self._h = None

    def _workspace(self, B: int, device) -> torch.Tensor:
        need = int(_native.lib().sv_prep_workspace_bytes(self._h, B))
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(max(need, 1), dtype=torch.uint8, device=device)
        return self._ws

    def images(self, frames_u8: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """uint8 [B, H, W, 3] CUDA -> float32 [B, 3, crop, crop] (Resize -> CenterCrop -> ToTensor -> Normalize)."""
        if not frames_u8.is_cuda:
            raise RuntimeError("FramePreprocessor runs on CUDA tensors only (no CPU fallback)")
        if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or tuple(frames_u8.shape[1:]) != (*self.in_hw, 3):
            raise ValueError(f"expected uint8 [B, {self.in_hw[0]}, {self.in_hw[1]}, 3], got {frames_u8.dtype} {tuple(frames_u8.shape)}")
        frames_u8 = frames_u8.contiguous()
        B = frames_u8.shape[0]
        if out is None:
            out = torch.empty((B, 3, self.crop, self.crop), dtype=torch.float32, device=frames_u8.device)
        ws = self._workspace(B, frames_u8.device)
        st = ctypes.c_void_p(torch.cuda.current_stream(frames_u8.device).cuda_stream)
        _native.check(_native.lib().sv_prep_images(self._h, ctypes.c_void_p(frames_u8.data_ptr()), B, ctypes.c_void_p(out.data_ptr()),
                                                   ctypes.c_void_p(ws.data_ptr()), ws.numel(), st), "sv_prep_images")
        return out

    def flow(self, flow_f32: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """float32 [B, Hf, Wf, 2] CUDA -> float32 [B, 2, crop, crop] (cv2.resize INTER_LINEAR, rescale, CenterCrop)."""
        if self.flow_hw is None:
            raise RuntimeError("this FramePreprocessor was created without a flow geometry")
        if not flow_f32.is_cuda:
            raise RuntimeError("FramePreprocessor runs on CUDA tensors only (no CPU fallback)")
        if flow_f32.dtype != torch.float32 or flow_f32.dim() != 4 or tuple(flow_f32.shape[1:]) != (*self.flow_hw, 2):
            raise ValueError(f"expected float32 [B, {self.flow_hw[0]}, {self.flow_hw[1]}, 2], got {flow_f32.dtype} {tuple(flow_f32.shape)}")
        flow_f32 = flow_f32.contiguous()
        B = flow_f32.shape[0]
        if out is None:
            out = torch.empty((B, 2, self.crop, self.crop), dtype=torch.float32, device=flow_f32.device)
        st = ctypes.c_void_p(torch.cuda.current_stream(flow_f32.device).cuda_stream)
        _native.check(_native.lib().sv_prep_flow(self._h, ctypes.c_void_p(flow_f32.data_ptr()), B, ctypes.c_void_p(out.data_ptr()), st), "sv_prep_flow")
        return out

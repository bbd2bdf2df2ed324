"""CPU restatement of the reference's per-frame input transforms.

TEST INFRASTRUCTURE ONLY: imported by tests/ and nothing else (the product path is csrc/preprocess.cu and refuses to run
without the CUDA library).

What the reference does to every frame before the encoder sees it:
  * image and segmentation map (both PIL RGB):  `transforms.Resize((250, 250))` -> `CenterCrop(224)` -> `ToTensor()` ->
    `Normalize(mean, std)`                                                   (generate_evp_LFB.py:242-248, crop_type 1)
  * RAFT flow (`.npy`, float32 [H, W, 2]):  `cv2.resize(flow, (250, 250), INTER_LINEAR)`, x/y displacement scaled by
    250/W and 250/H, `permute(2, 0, 1)`, then the geometric transforms only (`Resize((250,250))` is the identity on a
    250x250 tensor, `CenterCrop(224)`)                                        (data_process.py:424-481)

The arithmetic lives in third-party libraries that are not part of /root/reference:
  * Pillow (12.2.0 in this image) `Image.resize(size, BILINEAR)` — libImaging/Resample.c: separable, antialiased
    (support scaled by the downscale factor), horizontal pass first with a uint8 intermediate, coefficients rounded to
    22-bit fixed point, accumulators start at 2^21, results clamped to [0, 255].  Restated in `pil_bilinear_resize_u8`.
  * torchvision (0.26.0) `ToTensor` = uint8 -> float32, `/ 255`; `Normalize` = `(x - mean) / std` in float32;
    `CenterCrop` offset = int(round((250 - 224) / 2.0)) = 13.
  * OpenCV (4.13.0) `cv2.resize(..., INTER_LINEAR)` on float32: no antialiasing; source coordinate
    (d + 0.5) * scale - 0.5, floor, clamp at the borders, horizontal then vertical float32 lerp.  Restated in
    `cv2_linear_resize_f32` (OpenCV's SIMD path may contract a*b+c*d differently: parity with cv2 is to 1 ulp-ish, the
    test tolerance says so).

Pinned by tests/test_preprocess_cpu.py against the installed Pillow / torchvision / OpenCV themselves (bit-exact for the
uint8 image path).
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Resample.c
MEAN = (0.41757566, 0.26098573, 0.25888634)  # generate_evp_LFB.py:247
STD = (0.21938758, 0.1983, 0.19342837)


def pil_bilinear_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter (support 1.0) over the whole axis.
    Returns (bounds [out,2] int32 = (first source index, tap count), kk [out, ksize] int32 fixed point, ksize)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.zeros(ksize, dtype=np.float64)
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
        ww = w[:xmax].sum() if xmax > 0 else 0.0
        # Resample.c accumulates ww left to right in double; np.sum of <= a few dozen doubles may pair differently, so redo it serially
        ww = 0.0
        for x in range(xmax):
            ww += w[x]
        if ww != 0.0:
            w[:xmax] /= ww
        for x in range(ksize):
            kk[xx, x] = int(-0.5 + w[x] * (1 << PRECISION_BITS)) if w[x] < 0 else int(0.5 + w[x] * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _resample_axis_u8(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray, axis: int) -> np.ndarray:
    """One 8-bit pass of Resample.c (ImagingResampleHorizontal_8bpc / Vertical_8bpc): int32 accumulate from 2^21, >> 22, clamp."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + src.shape[1:], dtype=np.uint8)
    for o in range(bounds.shape[0]):
        x0, n = int(bounds[o, 0]), int(bounds[o, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for t in range(n):
            acc += src[x0 + t] * int(kk[o, t])
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_bilinear_resize_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """`PIL.Image.fromarray(img).resize((out_w, out_h), BILINEAR)` for uint8 [H, W, C].  Same-size resize is a copy
    (Image.resize returns self.copy()); otherwise horizontal pass (only if the width changes) then vertical pass."""
    H, W = img.shape[:2]
    out = img
    if W != out_w:
        b, k, _ = pil_bilinear_coeffs(W, out_w)
        out = _resample_axis_u8(out, b, k, axis=1)
    if H != out_h:
        b, k, _ = pil_bilinear_coeffs(H, out_h)
        out = _resample_axis_u8(out, b, k, axis=0)
    return out.copy() if out is img else out


def center_crop_offsets(h: int, w: int, crop: int) -> Tuple[int, int]:
    """torchvision.transforms.functional.center_crop: top = int(round((h - crop) / 2.0)), same for left."""
    return int(round((h - crop) / 2.0)), int(round((w - crop) / 2.0))


def image_transform(img_u8: np.ndarray, resize: int = 250, crop: int = 224, mean=MEAN, std=STD) -> np.ndarray:
    """uint8 [H, W, 3] -> float32 [3, crop, crop]: Resize((resize, resize)) -> CenterCrop(crop) -> ToTensor -> Normalize."""
    r = pil_bilinear_resize_u8(img_u8, resize, resize)
    top, left = center_crop_offsets(resize, resize, crop)
    c = r[top:top + crop, left:left + crop]
    x = c.astype(np.float32) / np.float32(255.0)                    # ToTensor: .to(float32).div(255)
    x = (x - np.asarray(mean, dtype=np.float32)) / np.asarray(std, dtype=np.float32)  # Normalize: sub_(mean).div_(std)
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def cv2_linear_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """OpenCV resize(INTER_LINEAR) source index and fraction per output index (imgproc/src/resize.cpp, the generic path):
    fx = (float)((d + 0.5) * scale - 0.5); s = floor(fx); fx -= s; s < 0 -> (0, 0); s >= in-1 -> (in-1, 0)."""
    scale = 1.0 / (out_size / in_size)  # resize.cpp: inv_scale_x = (double)dsize.width / ssize.width; scale_x = 1. / inv_scale_x
    idx = np.zeros(out_size, dtype=np.int32)
    frac = np.zeros(out_size, dtype=np.float32)
    for d in range(out_size):
        fx = np.float32((d + 0.5) * scale - 0.5)
        s = int(math.floor(fx))
        fx = np.float32(fx - np.float32(s))
        if s < 0:
            s, fx = 0, np.float32(0)
        if s >= in_size - 1:
            s, fx = in_size - 1, np.float32(0)
        idx[d], frac[d] = s, fx
    return idx, frac


def cv2_linear_resize_f32(a: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """cv2.resize(a, (out_w, out_h), interpolation=cv2.INTER_LINEAR) for float32 [H, W, C] (returns a copy when the size is unchanged)."""
    H, W = a.shape[:2]
    if (H, W) == (out_h, out_w):
        return a.copy()
    xi, xf = cv2_linear_coeffs(W, out_w)
    yi, yf = cv2_linear_coeffs(H, out_h)
    x1 = np.minimum(xi + 1, W - 1)
    y1 = np.minimum(yi + 1, H - 1)
    a0 = (np.float32(1) - xf)[None, :, None]
    a1 = xf[None, :, None]
    rows = a[:, xi, :] * a0 + a[:, x1, :] * a1                    # horizontal lerp of every source row, float32
    b0 = (np.float32(1) - yf)[:, None, None]
    b1 = yf[:, None, None]
    return (rows[yi] * b0 + rows[y1] * b1).astype(np.float32)      # vertical lerp


def flow_transform(flow: np.ndarray, resize: int = 250, crop: int = 224) -> np.ndarray:
    """float32 [H, W, 2] -> float32 [2, crop, crop] (data_process.py:432-447 + the CenterCrop applied at :461-480)."""
    H, W = flow.shape[:2]
    r = cv2_linear_resize_f32(flow, resize, resize)
    r[:, :, 0] *= np.float32(resize / W)   # numpy: float32 array *= python float -> float32 multiply
    r[:, :, 1] *= np.float32(resize / H)
    top, left = center_crop_offsets(resize, resize, crop)
    return np.ascontiguousarray(r[top:top + crop, left:left + crop].transpose(2, 0, 1))

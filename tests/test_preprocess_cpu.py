"""Pins oracle/preprocess_oracle.py to the third-party libraries whose arithmetic the reference's input transforms use
(Pillow / torchvision / OpenCV, all part of this image).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as po

PIL = pytest.importorskip("PIL.Image")
tvt = pytest.importorskip("torchvision.transforms")
cv2 = pytest.importorskip("cv2")

SIZES = [(480, 854), (250, 250), (256, 320), (1080, 1920), (200, 180), (250, 400)]  # (H, W): down-, identity, mixed, up-scale


def _img(h, w, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    base[: h // 4, : w // 4] = 255  # saturated block: exercises the clamp at the top of the range
    base[-(h // 5):, -(w // 5):] = 0
    return base


@pytest.mark.parametrize("hw", SIZES)
def test_pil_bilinear_restatement_is_bit_exact(hw):
    img = _img(*hw, seed=hw[0] * 7 + hw[1])
    want = np.asarray(PIL.fromarray(img, "RGB").resize((250, 250), PIL.BILINEAR))
    got = po.pil_bilinear_resize_u8(img, 250, 250)
    assert got.dtype == np.uint8 and got.shape == (250, 250, 3)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("hw", SIZES[:4])
def test_image_transform_matches_torchvision_bit_exact(hw):
    img = _img(*hw, seed=3)
    t = tvt.Compose([tvt.Resize((250, 250)), tvt.CenterCrop(224), tvt.ToTensor(), tvt.Normalize(list(po.MEAN), list(po.STD))])  # generate_evp_LFB.py:243-248
    want = t(PIL.fromarray(img, "RGB")).numpy()
    got = po.image_transform(img)
    assert got.dtype == np.float32 and got.shape == (3, 224, 224)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("hw", SIZES)
def test_flow_transform_matches_cv2_and_reference_steps(hw):
    h, w = hw
    rng = np.random.default_rng(h + w)
    flow = (rng.standard_normal((h, w, 2)) * 3.0).astype(np.float32)
    # data_process.py:432-447, then CenterCrop(224) on the [2,250,250] tensor (:461-480)
    r = cv2.resize(flow, (250, 250), interpolation=cv2.INTER_LINEAR)
    r[:, :, 0] *= 250 / w
    r[:, :, 1] *= 250 / h
    want = tvt.CenterCrop(224)(torch.from_numpy(r).permute(2, 0, 1).float()).numpy()
    got = po.flow_transform(flow)
    assert got.shape == (2, 224, 224)
    # OpenCV's vectorised lerp may fuse multiply-adds; the restatement rounds each product: agree to a few float32 ulps of the operands
    assert np.max(np.abs(got - want)) <= 4e-6 * max(1.0, float(np.max(np.abs(flow))))


def test_center_crop_offset_is_13():
    assert po.center_crop_offsets(250, 250, 224) == (13, 13)

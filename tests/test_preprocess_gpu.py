"""On-GPU input transforms (csrc/preprocess.cu) against the libraries the reference uses on the CPU and
against oracle/preprocess_oracle.py.  Image path: bit-exact.  Flow path: bit-exact vs the oracle, float32-rounding vs OpenCV."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SIZES = [(480, 854), (250, 250), (256, 320), (1080, 1920), (200, 180), (250, 400)]


def _frames(B, h, w, seed):
    rng = np.random.default_rng(seed)
    f = rng.integers(0, 256, size=(B, h, w, 3), dtype=np.uint8)
    f[:, : h // 4, : w // 4] = 255
    f[:, -(h // 5):, -(w // 5):] = 0
    return f


@pytest.mark.parametrize("hw", SIZES)
def test_images_bit_exact_vs_torchvision_and_oracle(hw):
    from PIL import Image
    import torchvision.transforms as tvt
    from oracle import preprocess_oracle as po
    from surgvid_b200.preprocess import FramePreprocessor
    h, w = hw
    B = 3
    frames = _frames(B, h, w, seed=h + 3 * w)
    pre = FramePreprocessor((h, w))
    got = pre.images(torch.from_numpy(frames).cuda()).cpu().numpy()
    assert got.shape == (B, 3, 224, 224) and got.dtype == np.float32
    t = tvt.Compose([tvt.Resize((250, 250)), tvt.CenterCrop(224), tvt.ToTensor(), tvt.Normalize(list(po.MEAN), list(po.STD))])  # generate_evp_LFB.py:243-248
    for b in range(B):
        want = t(Image.fromarray(frames[b], "RGB")).numpy()
        assert np.array_equal(got[b], want), f"frame {b}: {np.abs(got[b] - want).max()}"
        assert np.array_equal(got[b], po.image_transform(frames[b]))


@pytest.mark.parametrize("hw", SIZES)
def test_flow_vs_oracle_and_cv2(hw):
    import cv2
    import torchvision.transforms as tvt
    from oracle import preprocess_oracle as po
    from surgvid_b200.preprocess import FramePreprocessor
    h, w = hw
    B = 2
    rng = np.random.default_rng(h * 5 + w)
    flow = (rng.standard_normal((B, h, w, 2)) * 3.0).astype(np.float32)
    pre = FramePreprocessor((h, w), flow_hw=(h, w))
    got = pre.flow(torch.from_numpy(flow).cuda()).cpu().numpy()
    assert got.shape == (B, 2, 224, 224)
    for b in range(B):
        assert np.array_equal(got[b], po.flow_transform(flow[b]))
        r = cv2.resize(flow[b], (250, 250), interpolation=cv2.INTER_LINEAR)  # data_process.py:435-447
        r[:, :, 0] *= 250 / w
        r[:, :, 1] *= 250 / h
        want = tvt.CenterCrop(224)(torch.from_numpy(r).permute(2, 0, 1).float()).numpy()
        assert np.max(np.abs(got[b] - want)) <= 4e-6 * max(1.0, float(np.abs(flow[b]).max()))


def test_other_resize_and_crop_geometry():
    """crop_type 2 of the reference: Resize((224,224)) with no crop (generate_evp_LFB.py:250-254)."""
    from PIL import Image
    import torchvision.transforms as tvt
    from oracle import preprocess_oracle as po
    from surgvid_b200.preprocess import FramePreprocessor
    frames = _frames(2, 300, 410, seed=9)
    pre = FramePreprocessor((300, 410), resize=224, crop=224)
    got = pre.images(torch.from_numpy(frames).cuda()).cpu().numpy()
    t = tvt.Compose([tvt.Resize((224, 224)), tvt.ToTensor(), tvt.Normalize(list(po.MEAN), list(po.STD))])
    for b in range(2):
        assert np.array_equal(got[b], t(Image.fromarray(frames[b], "RGB")).numpy())


def test_rejects_cpu_tensors_and_wrong_shapes():
    from surgvid_b200.preprocess import FramePreprocessor
    pre = FramePreprocessor((250, 250))
    with pytest.raises(RuntimeError):
        pre.images(torch.zeros(1, 250, 250, 3, dtype=torch.uint8))
    with pytest.raises(ValueError):
        pre.images(torch.zeros(1, 240, 250, 3, dtype=torch.uint8, device="cuda"))
    with pytest.raises(RuntimeError):
        pre.flow(torch.zeros(1, 250, 250, 2, device="cuda"))


def test_raw_extraction_equals_preprocessed_extraction():
    """LFBExtractor.extract_raw(uint8 frames, uint8 segmaps, raw flow) == extract(transformed fp32 tensors): same kernels after the transforms."""
    from surgvid_b200 import synthetic
    from surgvid_b200.lfb import LFBExtractor
    from surgvid_b200.models.mix_transformer_evp import mit_b3_evp
    from surgvid_b200.preprocess import FramePreprocessor
    model = mit_b3_evp()
    model.load_state_dict(synthetic.synth_state_dict(synthetic.evp_key_shapes("mit_b3_evp"), seed=0, mode="stress"), strict=True)
    model = model.cuda().eval()
    N, h, w = 10, 250, 250
    frames = _frames(N, h, w, seed=1)
    segs = (np.random.default_rng(2).random((N, h, w, 1)) > 0.7).astype(np.uint8).repeat(3, axis=3) * 255
    flow = (np.random.default_rng(3).standard_normal((N, h, w, 2)) * 2.0).astype(np.float32)
    ex = LFBExtractor(model, batch_size=4)
    got = ex.extract_raw(torch.from_numpy(frames), torch.from_numpy(segs), torch.from_numpy(flow)).clone()
    raw_h2d = ex.h2d_bytes
    pre = FramePreprocessor((h, w), flow_hw=(h, w))
    x = pre.images(torch.from_numpy(frames).cuda()).cpu()
    s = pre.images(torch.from_numpy(segs).cuda()).cpu()
    f = pre.flow(torch.from_numpy(flow).cuda()).cpu()
    want = ex.extract(x, s, f)
    assert torch.equal(got, want)
    assert raw_h2d == N * (2 * h * w * 3 + h * w * 2 * 4)


def test_extract_videos_pipelined_equals_per_video_calls():
    """LFBExtractor.extract_videos (one pipelined pass over several videos, only the first ramped up) == extract() per video."""
    from surgvid_b200 import synthetic
    from surgvid_b200.lfb import LFBExtractor
    from surgvid_b200.models.mix_transformer_evp import mit_b3_evp
    model = mit_b3_evp()
    model.load_state_dict(synthetic.synth_state_dict(synthetic.evp_key_shapes("mit_b3_evp"), seed=0, mode="stress"), strict=True)
    model = model.cuda().eval()
    vids = []
    for i, n in enumerate((9, 4, 13)):
        x, s, f = synthetic.synth_frames(n, seed=300 + i)
        vids.append((x.pin_memory(), s.pin_memory(), f.pin_memory()))
    ex = LFBExtractor(model, batch_size=4)
    got = [t.clone() for t in ex.extract_videos(vids)]
    assert ex.h2d_bytes == sum(v[0].shape[0] for v in vids) * 8 * 224 * 224 * 4
    for v, g in zip(vids, got):
        assert torch.equal(g, ex.extract(*v))
    with pytest.raises(ValueError):
        ex.extract_videos([vids[0], (vids[1][0], vids[1][1], None)])

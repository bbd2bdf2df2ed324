"""LFB pickle / phase-file formats: byte-level agreement with what the reference drivers write and read."""
import pickle

import numpy as np
import pytest

import surgvid_b200  # noqa: F401
from surgvid_b200 import lfb_io


def test_pickle_matches_reference_writer(tmp_path):
    rng = np.random.default_rng(0)
    blocks = [rng.standard_normal((n, 2048)).astype(np.float32) for n in (200, 200, 37)]
    # what generate_evp_LFB.py does: float64 seed array + np.concatenate per batch, then pickle.dump (:295-297, 457, 513-520)
    g = np.zeros(shape=(0, 2048))
    for b in blocks:
        g = np.concatenate((g, b), axis=0)
    ref_path = tmp_path / "ref.pkl"
    with open(ref_path, "wb") as f:
        pickle.dump(g, f)
    bank = lfb_io.LFBBank(437)
    for b in blocks:
        bank.append(b)
    (mine,) = lfb_io.save_lfb_pickles(str(tmp_path / "LFB1"), {"val": bank.array()})
    assert mine.endswith("evp_LFB_val.pkl")
    a, b = lfb_io.load_lfb_pickle(str(ref_path)), lfb_io.load_lfb_pickle(mine)
    assert a.dtype == b.dtype == np.float64 and a.shape == b.shape == (437, 2048)
    assert np.array_equal(a, b)
    assert open(ref_path, "rb").read() == open(mine, "rb").read()   # identical bytes on disk
    with pytest.raises(ValueError):
        lfb_io.LFBBank(10).array()
    with pytest.raises(ValueError):
        bank.append(blocks[0])


def test_video_slicing_matches_get_long_feature():
    lfb = np.arange(10 * 4, dtype=np.float64).reshape(10, 4)
    num_each = [3, 5, 2]
    sl = lfb_io.video_slices(num_each)
    assert [(s.start, s.stop) for s in sl] == [(0, 3), (3, 8), (8, 10)]
    # reference: list of rows wrapped in a list -> np.array(...) of shape [1, T, dim]
    ref = np.array([[lfb[3 + k] for k in range(5)]])
    assert np.array_equal(lfb_io.long_feature(lfb, 3, 5), ref)


def test_phase_file_format(tmp_path):
    p = tmp_path / "video41-phase.txt"
    phases = [0, 0, 1, 3, 6]
    lfb_io.write_phase_file(str(p), phases)
    # reference writer (trans_SV_output.py:313-320)
    expect = "".join(str(c * 25) + "\t" + str(ph) + "\t" + "\n" for c, ph in enumerate(phases))
    assert open(p).read() == expect
    back = lfb_io.read_phase_file(str(p))
    assert back[:, 0].tolist() == [0, 25, 50, 75, 100] and back[:, 1].tolist() == phases

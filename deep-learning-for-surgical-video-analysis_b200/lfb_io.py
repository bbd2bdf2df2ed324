"""LFB on-disk format and phase-file writer — the data formats either side of the hot path.

The reference writes three pickles of a float64 ndarray `[N_frames, 2048]` (generate_evp_LFB.py:295-297, 457, 513-520:
`evp_LFB_train.pkl`, `evp_LFB_val.pkl`, `evp_LFB_test.pkl`; rows in video order) which `trans_SV_output.py:106-111` /
`tecno*.py` read back and slice per video by cumulative frame counts (`get_long_feature`, trans_SV_output.py:78-87,
268-271).  Predictions are written one file per video, `video<ID>-phase.txt`, lines `"<frame*fps>\\t<phase>\\t\\n"` with
fps = 25 (trans_SV_output.py:304-321).  Same bytes-on-disk contract here, without the reference's O(N^2) `np.concatenate`.
"""
from __future__ import annotations

import os
import pickle
from typing import Dict, Iterable, List, Sequence

import numpy as np

LFB_FILENAMES = {"train": "evp_LFB_train.pkl", "val": "evp_LFB_val.pkl", "test": "evp_LFB_test.pkl"}


class LFBBank:
    """Preallocated float64 `[N, dim]` bank filled block by block in video order."""

    def __init__(self, total_frames: int, dim: int = 2048):
        self.data = np.zeros((int(total_frames), dim), dtype=np.float64)
        self.filled = 0

    def append(self, feats) -> None:
        a = np.asarray(feats.detach().cpu().numpy() if hasattr(feats, "detach") else feats)
        n = a.shape[0]
        if self.filled + n > self.data.shape[0] or a.shape[1] != self.data.shape[1]:
            raise ValueError(f"LFB block of shape {a.shape} does not fit at row {self.filled} of {self.data.shape}")
        self.data[self.filled:self.filled + n] = a  # fp32 -> float64 is exact
        self.filled += n

    def array(self) -> np.ndarray:
        if self.filled != self.data.shape[0]:
            raise ValueError(f"LFB bank incomplete: {self.filled} of {self.data.shape[0]} rows filled")
        return self.data


def save_lfb_pickles(save_dir: str, banks: Dict[str, np.ndarray]) -> List[str]:
    """Write `evp_LFB_{train,val,test}.pkl` exactly as generate_evp_LFB.py:513-520 does (pickle of a float64 ndarray)."""
    os.makedirs(save_dir, exist_ok=True)
    paths = []
    for split, arr in banks.items():
        if split not in LFB_FILENAMES:
            raise KeyError(f"unknown split '{split}' (expected one of {sorted(LFB_FILENAMES)})")
        arr = np.asarray(arr)
        if arr.ndim != 2:
            raise ValueError("LFB array must be [N_frames, dim]")
        path = os.path.join(save_dir, LFB_FILENAMES[split])
        with open(path, "wb") as f:
            pickle.dump(np.ascontiguousarray(arr, dtype=np.float64), f)
        paths.append(path)
    return paths


def load_lfb_pickle(path: str) -> np.ndarray:
    """Read a reference-written (or our) LFB pickle: float64 ndarray [N_frames, dim] (trans_SV_output.py:106-111)."""
    with open(path, "rb") as f:
        arr = pickle.load(f)
    arr = np.asarray(arr)
    if arr.ndim != 2:
        raise ValueError(f"{path}: expected a 2-D feature array, got shape {arr.shape}")
    return arr


def video_slices(num_each: Sequence[int]) -> List[slice]:
    """Row ranges of each video inside a split's bank (cumulative `num_each`, trans_SV_output.py:56-72)."""
    out, start = [], 0
    for n in num_each:
        out.append(slice(start, start + int(n)))
        start += int(n)
    return out


def long_feature(lfb: np.ndarray, start_index: int, length: int) -> np.ndarray:
    """`get_long_feature` (trans_SV_output.py:78-87) as a view: [1, T, dim] rows of one video."""
    return lfb[int(start_index):int(start_index) + int(length)][None]


def write_phase_file(path: str, phases: Iterable[int], fps: int = 25) -> None:
    """`videoNN-phase.txt` as written at trans_SV_output.py:304-321: '<cnt*fps>\\t<phase>\\t\\n' per frame."""
    with open(path, "w") as f:
        for cnt, ph in enumerate(phases):
            f.write(str(cnt * fps) + "\t")
            f.write(str(int(ph)) + "\t")
            f.write("\n")


def read_phase_file(path: str) -> np.ndarray:
    rows = []
    with open(path) as f:
        for line in f:
            parts = line.rstrip("\n").split("\t")
            if len(parts) >= 2 and parts[0] != "":
                rows.append((int(parts[0]), int(parts[1])))
    return np.asarray(rows, dtype=np.int64).reshape(-1, 2)

"""Host-side mirror of the reference's LFB driver loop (generate_evp_LFB.py:439-499) and of the per-video MS-TCN
call (trans_SV_output.py:250-301), plus the shard-by-video logic for multi-GPU runs.

Frames and videos are independent, so N GPUs means N processes each extracting its own videos — there is no
collective on the data path.  What the reference does per batch, and what happens here instead:
  reference: `.to(device)` of pageable fp32 tensors, forward, `.cpu().numpy()`, `np.concatenate` (O(N^2) host copies,
             float64 result because the seed array is float64, generate_evp_LFB.py:295-297,457);
  here:      pinned host staging, H2D copies on a side stream overlapped with the previous batch's kernels, D2H of the
             [B,2048] features into a preallocated pinned block; the float64 view is produced once at the end.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch


def bind_to_gpu_numa_node(device_index: int) -> Optional[List[int]]:
    """Restrict this process to the CPU cores that are local to GPU `device_index` (NVML's CPU affinity for the device), so that the
    pinned staging buffers it allocates afterwards are first-touched on that NUMA node and every rank's H2D stream reads local memory.
    With 8 ranks feeding 8 GPUs from one socket's memory the host side, not PCIe, limits the end-to-end rate.  Returns the core
    list, or None when NVML / the affinity call is unavailable (nothing is changed then)."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [w * 64 + b for w, mask in enumerate(words) for b in range(64) if (mask >> b) & 1 and w * 64 + b < n_cpu]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


def lpt_assign(lengths: Sequence[int], n_ranks: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of videos to ranks. Returns per-rank video indices, ascending."""
    loads = [0] * n_ranks
    buckets: List[List[int]] = [[] for _ in range(n_ranks)]
    for v in sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i)):
        r = min(range(n_ranks), key=lambda j: (loads[j], j))
        buckets[r].append(v)
        loads[r] += int(lengths[v])
    return [sorted(b) for b in buckets]


def gather_in_video_order(per_rank_blocks: Sequence[Sequence[np.ndarray]], assignment: Sequence[Sequence[int]], n_videos: int) -> np.ndarray:
    """Host-side ordered gather: rank r produced one [T_v, D] block per video in `assignment[r]`; the LFB file is the
    concatenation by video index (generate_evp_LFB.py:457; consumers slice by cumulative num_each, trans_SV_output.py:56-72)."""
    slots: List[Optional[np.ndarray]] = [None] * n_videos
    for blocks, vids in zip(per_rank_blocks, assignment):
        if len(blocks) != len(vids):
            raise ValueError("a rank returned a different number of blocks than it was assigned videos")
        for blk, v in zip(blocks, vids):
            slots[v] = blk
    if any(s is None for s in slots):
        raise ValueError("missing feature block for video(s) " + str([i for i, s in enumerate(slots) if s is None]))
    return np.concatenate(slots, axis=0)


def ramp_schedule(n_frames: int, batch_size: int, ramp_start: int) -> List[tuple]:
    """(first frame, count) of the batches of one call: small first batches so the kernels start while the bulk of the input is
    still crossing PCIe (only the first copy of a call is not overlapped with compute), doubling up to batch_size."""
    starts, b0, ramp = [], 0, max(1, int(ramp_start))
    while b0 < n_frames:
        n = min(ramp, batch_size, n_frames - b0)
        starts.append((b0, n))
        b0 += n
        ramp *= 2
    return starts


class LFBExtractor:
    """End-to-end feature extraction from HOST buffers through the drop-in model (the call a user of the reference
    makes, with the reference's batch size of 200 by default: generate_evp_LFB.py:36 `--val`)."""

    def __init__(self, model, batch_size: int = 200, device: Optional[torch.device] = None, ramp_start: Optional[int] = None):
        self.model = model
        self.batch_size = int(batch_size)
        # first batch of the ramp-up schedule (see _schedule); default batch_size / 8 (measured best of 25..800 at batch 800, scripts/e2e_ramp.py)
        self.ramp_start = max(1, int(ramp_start) if ramp_start is not None else self.batch_size // 8)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._copy_stream = torch.cuda.Stream(self.device)
        self._dev = None  # double-buffered device staging
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _schedule(self, N: int):
        return ramp_schedule(N, self.batch_size, self.ramp_start)

    def _staging(self, H, W, with_flow):
        key = (H, W, with_flow)
        if self._dev is None or self._dev[0] != key:
            bufs = []
            for _ in range(2):
                x = torch.empty((self.batch_size, 3, H, W), dtype=torch.float32, device=self.device)
                s = torch.empty((self.batch_size, 3, H, W), dtype=torch.float32, device=self.device)
                f = torch.empty((self.batch_size, 2, H, W), dtype=torch.float32, device=self.device) if with_flow else None
                bufs.append((x, s, f, torch.cuda.Event(), torch.cuda.Event()))
            self._dev = (key, bufs)
        return self._dev[1]

    @torch.no_grad()
    def extract(self, frames: torch.Tensor, segmaps: torch.Tensor, flow: Optional[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """frames/segmaps: [N,(1,)3,H,W] fp32 HOST (pinned for full overlap), flow: [N,(1,)2,H,W] or None
        -> [N, 2048] fp32 pinned host tensor (features in input order)."""
        return self.extract_videos([(frames, segmaps, flow)], outs=None if out is None else [out])[0]

    @torch.no_grad()
    def extract_videos(self, videos: Sequence, outs: Optional[Sequence[torch.Tensor]] = None) -> List[torch.Tensor]:
        """The reference driver's loop over videos (generate_evp_LFB.py:439-499) as ONE pipelined pass: `videos` is a sequence of
        (frames, segmaps, flow-or-None) HOST tensors shaped as for `extract` (same H x W and flow presence for all); the copy of
        every batch overlaps the kernels of the previous one across video boundaries, so only the very first batch of the call
        is exposed (and only the first video is ramped up).  Returns one [T_v, 2048] fp32 pinned host tensor per video."""
        if len(videos) == 0:
            return []
        H, W = videos[0][0].shape[-2], videos[0][0].shape[-1]
        with_flow = videos[0][2] is not None
        vids = []
        for (fr, sg, fl) in videos:
            N = fr.shape[0]
            if fr.shape[-2:] != (H, W) or (fl is not None) != with_flow:
                raise ValueError("extract_videos: all videos of a call must share H x W and the presence of flow")
            vids.append((N, fr.reshape(N, 3, H, W), sg.reshape(N, 3, H, W), None if fl is None else fl.reshape(N, 2, H, W)))
        bufs = self._staging(H, W, with_flow)
        D = self.model.embedding_dim
        if outs is None:
            outs = [torch.empty((v[0], D), dtype=torch.float32).pin_memory() for v in vids]
        compute = torch.cuda.current_stream(self.device)
        self.h2d_bytes = self.d2h_bytes = 0
        batches = []
        for vi, v in enumerate(vids):
            sched = self._schedule(v[0]) if vi == 0 else [(b0, min(self.batch_size, v[0] - b0)) for b0 in range(0, v[0], self.batch_size)]
            batches += [(vi, b0, n) for (b0, n) in sched]
        for bi, (vi, b0, n) in enumerate(batches):
            _, frames, segmaps, flow = vids[vi]
            x, s, f, ev_in, ev_free = bufs[bi % 2]
            with torch.cuda.stream(self._copy_stream):
                if bi >= 2:
                    self._copy_stream.wait_event(ev_free)  # kernels that read this staging buffer have finished
                x[:n].copy_(frames[b0:b0 + n], non_blocking=True)
                s[:n].copy_(segmaps[b0:b0 + n], non_blocking=True)
                self.h2d_bytes += 2 * n * 3 * H * W * 4
                if f is not None:
                    f[:n].copy_(flow[b0:b0 + n], non_blocking=True)
                    self.h2d_bytes += n * 2 * H * W * 4
                ev_in.record(self._copy_stream)
            compute.wait_event(ev_in)
            feats = self.model(x[:n], s[:n], None if f is None else f[:n], return_features=True)
            ev_free.record(compute)
            outs[vi][b0:b0 + n].copy_(feats, non_blocking=True)
            self.d2h_bytes += n * D * 4
        compute.synchronize()
        return list(outs)

    @torch.no_grad()
    def extract_raw(self, frames_u8: torch.Tensor, segmaps_u8: torch.Tensor, flow_raw: Optional[torch.Tensor], resize: int = 250,
                    crop: int = 224, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Same as `extract`, but from what the reference's dataset class starts from after decoding:
        frames/segmaps uint8 [N, H, W, 3] HOST, flow float32 [N, Hf, Wf, 2] HOST (raw RAFT field) or None.  The
        Resize/CenterCrop/ToTensor/Normalize and the flow resize+rescale run on the GPU (preprocess.FramePreprocessor), so
        uint8 frames cross PCIe instead of normalised fp32."""
        return self.extract_raw_videos([(frames_u8, segmaps_u8, flow_raw)], resize=resize, crop=crop, outs=None if out is None else [out])[0]

    @torch.no_grad()
    def extract_raw_videos(self, videos: Sequence, resize: int = 250, crop: int = 224,
                           outs: Optional[Sequence[torch.Tensor]] = None) -> List[torch.Tensor]:
        """`extract_videos` for raw inputs: `videos` is a sequence of (frames_u8, segmaps_u8, flow_raw-or-None) HOST tensors (one frame
        size and one flow size per call); one pipelined pass, only the first video ramped up."""
        from .preprocess import FramePreprocessor
        if len(videos) == 0:
            return []
        H, W = videos[0][0].shape[1], videos[0][0].shape[2]
        fhw = None if videos[0][2] is None else (videos[0][2].shape[1], videos[0][2].shape[2])
        for (fr, sg, fl) in videos:
            if (fr.shape[1], fr.shape[2]) != (H, W) or (None if fl is None else (fl.shape[1], fl.shape[2])) != fhw:
                raise ValueError("extract_raw_videos: all videos of a call must share the frame size and the flow size")
        key = (H, W, fhw, resize, crop)
        if getattr(self, "_prep_key", None) != key:
            self._prep = FramePreprocessor((H, W), flow_hw=fhw, resize=resize, crop=crop)
            self._prep_key = key
            self._raw = []
            for _ in range(2):
                fu = torch.empty((self.batch_size, H, W, 3), dtype=torch.uint8, device=self.device)
                su = torch.empty((self.batch_size, H, W, 3), dtype=torch.uint8, device=self.device)
                fl = None if fhw is None else torch.empty((self.batch_size, fhw[0], fhw[1], 2), dtype=torch.float32, device=self.device)
                self._raw.append((fu, su, fl, torch.cuda.Event(), torch.cuda.Event()))
            self._pre_out = (torch.empty((self.batch_size, 3, crop, crop), dtype=torch.float32, device=self.device),
                             torch.empty((self.batch_size, 3, crop, crop), dtype=torch.float32, device=self.device),
                             None if fhw is None else torch.empty((self.batch_size, 2, crop, crop), dtype=torch.float32, device=self.device))
        D = self.model.embedding_dim
        if outs is None:
            outs = [torch.empty((v[0].shape[0], D), dtype=torch.float32).pin_memory() for v in videos]
        compute = torch.cuda.current_stream(self.device)
        self.h2d_bytes = self.d2h_bytes = 0
        batches = []
        for vi, v in enumerate(videos):
            N = v[0].shape[0]
            sched = self._schedule(N) if vi == 0 else [(b0, min(self.batch_size, N - b0)) for b0 in range(0, N, self.batch_size)]
            batches += [(vi, b0, n) for (b0, n) in sched]
        x, s, f = self._pre_out
        for bi, (vi, b0, n) in enumerate(batches):
            frames_u8, segmaps_u8, flow_raw = videos[vi]
            fu, su, fl, ev_in, ev_free = self._raw[bi % 2]
            with torch.cuda.stream(self._copy_stream):
                if bi >= 2:
                    self._copy_stream.wait_event(ev_free)
                fu[:n].copy_(frames_u8[b0:b0 + n], non_blocking=True)
                su[:n].copy_(segmaps_u8[b0:b0 + n], non_blocking=True)
                self.h2d_bytes += 2 * n * H * W * 3
                if fl is not None:
                    fl[:n].copy_(flow_raw[b0:b0 + n], non_blocking=True)
                    self.h2d_bytes += n * fhw[0] * fhw[1] * 2 * 4
                ev_in.record(self._copy_stream)
            compute.wait_event(ev_in)
            self._prep.images(fu[:n], out=x[:n])
            self._prep.images(su[:n], out=s[:n])
            if fl is not None:
                self._prep.flow(fl[:n], out=f[:n])
            ev_free.record(compute)  # the raw staging buffers are free once the transforms have run
            feats = self.model(x[:n], s[:n], None if fl is None else f[:n], return_features=True)
            outs[vi][b0:b0 + n].copy_(feats, non_blocking=True)
            self.d2h_bytes += n * D * 4
        compute.synchronize()
        return list(outs)

    def extract_float64(self, frames, segmaps, flow) -> np.ndarray:
        """Same values as `extract`, as the float64 ndarray the reference pickles (generate_evp_LFB.py:513-520)."""
        return self.extract(frames, segmaps, flow).numpy().astype(np.float64)


@torch.no_grad()
def run_mstcn_per_video(mstcn_model, lfb: torch.Tensor, lengths: Sequence[int]):
    """The reference's per-video loop (trans_SV_output.py:250-301, MS-TCN part) done as ONE batched launch sequence:
    returns the last stage's logits per video, each [out_features, T_v] (== model.forward(video_fe)[-1].squeeze(1)[0])."""
    logits = mstcn_model.forward_videos(lfb, lengths)  # [stages, C, T_total]
    outs, o = [], 0
    for T in lengths:
        outs.append(logits[-1, :, o:o + T])
        o += T
    return outs

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


class SharedLFB:
    """The LFB array of a whole run — [sum(lengths), dim] rows in VIDEO order (generate_evp_LFB.py:457; consumers slice it by
    cumulative `num_each`, trans_SV_output.py:56-72) — in POSIX shared memory, mapped by every rank of the box and page-locked
    in each of them.  Every rank's device-to-host copies land directly on the rows of its own videos, so the "host-side gather
    of feature blocks in video order" is complete when the last rank's copy stream drains: no collective, no second host copy.
    Rank 0 creates (`create=True`), the others attach after a barrier; `unlink()` on rank 0 when done."""

    def __init__(self, name: str, lengths: Sequence[int], dim: int = 2048, create: bool = False, dtype: torch.dtype = torch.float32,
                 directory: Optional[str] = None, pin: bool = True):
        import os
        self.lengths = [int(n) for n in lengths]
        self.offsets = np.zeros(len(self.lengths) + 1, dtype=np.int64)
        self.offsets[1:] = np.cumsum(self.lengths)
        self.rows, self.dim = int(self.offsets[-1]), int(dim)
        itemsize = torch.empty((), dtype=dtype).element_size()
        self.nbytes = self.rows * self.dim * itemsize
        if directory is None:
            directory = "/dev/shm"
            try:
                st = os.statvfs(directory)
                if create and st.f_bavail * st.f_frsize < self.nbytes + (64 << 20):
                    directory = "/tmp"   # a page-cache backed MAP_SHARED file works the same way
            except OSError:
                directory = "/tmp"
        self.path = name if os.path.isabs(name) else os.path.join(directory, name)
        if create:
            with open(self.path, "wb") as f:
                f.truncate(self.nbytes)
        elif not os.path.exists(self.path) and not os.path.isabs(name):
            alt = os.path.join("/tmp", name)   # the creator fell back to /tmp
            if os.path.exists(alt):
                self.path = alt
        if os.path.getsize(self.path) != self.nbytes:
            raise ValueError(f"SharedLFB: {self.path} has {os.path.getsize(self.path)} bytes, expected {self.nbytes}")
        self.array = torch.from_file(self.path, shared=True, size=self.rows * self.dim, dtype=dtype).view(self.rows, self.dim)
        self.creator = bool(create)
        self.pinned = False
        if pin and torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.array.data_ptr(), self.nbytes, 0)
            self.pinned = int(rc) == 0   # unpinned still works (the copies become synchronous)

    def block(self, video: int) -> torch.Tensor:
        """Rows of one video: a [T_v, dim] view of the shared array."""
        return self.array[int(self.offsets[video]):int(self.offsets[video + 1])]

    def blocks(self, videos: Sequence[int]) -> List[torch.Tensor]:
        return [self.block(v) for v in videos]

    def close(self):
        if self.pinned:
            torch.cuda.cudart().cudaHostUnregister(self.array.data_ptr())
            self.pinned = False

    def unlink(self):
        import os
        self.close()
        if self.creator and os.path.exists(self.path):
            os.unlink(self.path)


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


class CyclicFrames:
    """A video of `length` frames whose frame t is row (start + t) % P of a P-frame host (or device) pool: lets a synthetic
    80-video workload (184 578 frames, 295 GB as fp32 tensors) be fed from a few GB of pinned memory.  Quacks like the
    [N, ...] tensor `LFBExtractor` slices: `.shape` and `.pieces(b0, n)` (contiguous pool slices covering frames b0..b0+n)."""

    def __init__(self, pool: torch.Tensor, start: int, length: int):
        self.pool, self.start, self.length = pool, int(start) % pool.shape[0], int(length)
        self.shape = (self.length,) + tuple(pool.shape[1:])

    def pieces(self, b0: int, n: int):
        P = self.pool.shape[0]
        a = (self.start + b0) % P
        while n > 0:
            m = min(n, P - a)
            yield self.pool[a:a + m]
            n -= m
            a = (a + m) % P


def _as_frames(t, trailing):
    """[N,(1,)C,H,W]-like tensor -> [N, *trailing] view; CyclicFrames pass through (their pool already has that layout)."""
    if isinstance(t, CyclicFrames):
        if tuple(t.shape[1:]) != tuple(trailing):   # e.g. a [P,1,3,H,W] pool (the reference's sequence_length-1 axis): view it as [P,3,H,W]
            return CyclicFrames(t.pool.reshape((t.pool.shape[0],) + tuple(trailing)), t.start, t.length)
        return t
    return t.reshape((t.shape[0],) + tuple(trailing))


def _pieces(t, b0, n):
    if isinstance(t, CyclicFrames):
        yield from t.pieces(b0, n)
    else:
        yield t[b0:b0 + n]


def pack_batches(lengths: Sequence[int], batch_size: int, ramp_start: Optional[int] = None) -> List[List[tuple]]:
    """The frames of all videos of a call as ONE stream cut into batches of <= batch_size frames; a batch is a list of
    (video, first frame, count) segments and may cross video boundaries (frames are independent in the encoder, so ragged
    tails at every video end would only waste launches).  With `ramp_start`, the first batches of the call are small
    (ramp_start, 2*ramp_start, ...) so the kernels start while the bulk of the input is still crossing PCIe."""
    batches: List[List[tuple]] = []
    cap = batch_size if not ramp_start else max(1, min(int(ramp_start), batch_size))
    cur: List[tuple] = []
    room = cap
    for vi, T in enumerate(lengths):
        b0 = 0
        while b0 < T:
            n = min(room, T - b0)
            cur.append((vi, b0, n))
            b0 += n
            room -= n
            if room == 0:
                batches.append(cur)
                cur = []
                cap = min(batch_size, cap * 2)
                room = cap
    if cur:
        batches.append(cur)
    return batches


class LFBExtractor:
    """End-to-end feature extraction from HOST buffers through the drop-in model (the call a user of the reference
    makes, with the reference's batch size of 200 by default: generate_evp_LFB.py:36 `--val`)."""

    def __init__(self, model, batch_size: int = 200, device: Optional[torch.device] = None, ramp_start: Optional[int] = None):
        self.model = model
        self.batch_size = int(batch_size)
        # first batch of the ramp-up schedule (see pack_batches); default batch_size / 8 (measured best of 25..800 at batch 800, scripts/e2e_ramp.py)
        self.ramp_start = max(1, int(ramp_start) if ramp_start is not None else self.batch_size // 8)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._copy_stream = torch.cuda.Stream(self.device)
        self._dev = None  # double-buffered device staging
        self._raw = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _schedule(self, N: int):
        return ramp_schedule(N, self.batch_size, self.ramp_start)

    def _fresh_buffers(self, old):
        """Called around (re)allocation of staging buffers: copies still in flight on the side stream may reference the old
        ones, and the caching allocator may hand out blocks that pending compute-stream work still uses."""
        compute = torch.cuda.current_stream(self.device)
        if old is not None:
            self._copy_stream.synchronize()
            compute.synchronize()
        self._copy_stream.wait_stream(compute)

    def _staging(self, H, W, with_flow):
        key = (H, W, with_flow)
        if self._dev is None or self._dev[0] != key:
            old = self._dev
            self._dev = None
            self._fresh_buffers(old)
            del old
            bufs = []
            for _ in range(2):
                x = torch.empty((self.batch_size, 3, H, W), dtype=torch.float32, device=self.device)
                s = torch.empty((self.batch_size, 3, H, W), dtype=torch.float32, device=self.device)
                f = torch.empty((self.batch_size, 2, H, W), dtype=torch.float32, device=self.device) if with_flow else None
                bufs.append((x, s, f, torch.cuda.Event(), torch.cuda.Event()))
            self._copy_stream.wait_stream(torch.cuda.current_stream(self.device))
            self._dev = (key, bufs)
        return self._dev[1]

    @torch.no_grad()
    def extract(self, frames: torch.Tensor, segmaps: torch.Tensor, flow: Optional[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """frames/segmaps: [N,(1,)3,H,W] fp32 HOST (pinned for full overlap), flow: [N,(1,)2,H,W] or None
        -> [N, 2048] fp32 pinned host tensor (features in input order)."""
        return self.extract_videos([(frames, segmaps, flow)], outs=None if out is None else [out])[0]

    @torch.no_grad()
    def extract_videos(self, videos: Sequence, outs: Optional[Sequence[torch.Tensor]] = None,
                       device_outs: Optional[Sequence[torch.Tensor]] = None) -> List[torch.Tensor]:
        """The reference driver's loop over videos (generate_evp_LFB.py:439-499) as ONE pipelined pass: `videos` is a sequence of
        (frames, segmaps, flow-or-None) HOST tensors shaped as for `extract` (same H x W and flow presence for all; `CyclicFrames`
        are accepted).  The frames of all videos form one stream cut into batches of `batch_size` that may cross video
        boundaries; the copy of every batch overlaps the kernels of the previous one, so only the very first (ramped-up)
        batch of the call is exposed.  Returns one [T_v, 2048] fp32 host tensor per video (`outs` if given — e.g. the rows of
        a `SharedLFB`); `device_outs` (one [T_v, 2048] CUDA tensor per video) additionally keeps the features on the GPU for
        the MS-TCN pass that follows (trans_SV_output.py:268-280)."""
        if len(videos) == 0:
            return []
        H, W = videos[0][0].shape[-2], videos[0][0].shape[-1]
        with_flow = videos[0][2] is not None
        vids = []
        for (fr, sg, fl) in videos:
            N = fr.shape[0]
            if tuple(fr.shape[-2:]) != (H, W) or (fl is not None) != with_flow:
                raise ValueError("extract_videos: all videos of a call must share H x W and the presence of flow")
            vids.append((N, _as_frames(fr, (3, H, W)), _as_frames(sg, (3, H, W)), None if fl is None else _as_frames(fl, (2, H, W))))
        bufs = self._staging(H, W, with_flow)
        D = self.model.embedding_dim
        if outs is None:
            outs = [torch.empty((v[0], D), dtype=torch.float32).pin_memory() for v in vids]
        compute = torch.cuda.current_stream(self.device)
        self.h2d_bytes = self.d2h_bytes = 0
        batches = pack_batches([v[0] for v in vids], self.batch_size, self.ramp_start)
        for bi, segs in enumerate(batches):
            x, s, f, ev_in, ev_free = bufs[bi % 2]
            with torch.cuda.stream(self._copy_stream):
                if bi >= 2:
                    self._copy_stream.wait_event(ev_free)  # kernels that read this staging buffer have finished
                o = 0
                for (vi, b0, n) in segs:
                    _, frames, segmaps, flow = vids[vi]
                    for dst, src in ((x, frames), (s, segmaps)) + (((f, flow),) if f is not None else ()):
                        oo = o
                        for piece in _pieces(src, b0, n):
                            dst[oo:oo + piece.shape[0]].copy_(piece, non_blocking=True)
                            oo += piece.shape[0]
                    o += n
                self.h2d_bytes += o * (6 + (2 if f is not None else 0)) * H * W * 4
                ev_in.record(self._copy_stream)
            compute.wait_event(ev_in)
            feats = self.model(x[:o], s[:o], None if f is None else f[:o], return_features=True)
            ev_free.record(compute)
            o = 0
            for (vi, b0, n) in segs:
                outs[vi][b0:b0 + n].copy_(feats[o:o + n], non_blocking=True)
                if device_outs is not None:
                    device_outs[vi][b0:b0 + n].copy_(feats[o:o + n], non_blocking=True)
                o += n
            self.d2h_bytes += o * D * 4
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
    def extract_raw_videos(self, videos: Sequence, resize: int = 250, crop: int = 224, outs: Optional[Sequence[torch.Tensor]] = None,
                           device_outs: Optional[Sequence[torch.Tensor]] = None) -> List[torch.Tensor]:
        """`extract_videos` for raw inputs: `videos` is a sequence of (frames_u8, segmaps_u8, flow_raw-or-None) HOST tensors (one frame
        size and one flow size per call; `CyclicFrames` accepted); one pipelined pass over the frame stream of all videos."""
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
            old = self._raw
            self._raw = None
            self._fresh_buffers(old)
            del old
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
            self._copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        D = self.model.embedding_dim
        if outs is None:
            outs = [torch.empty((v[0].shape[0], D), dtype=torch.float32).pin_memory() for v in videos]
        compute = torch.cuda.current_stream(self.device)
        self.h2d_bytes = self.d2h_bytes = 0
        batches = pack_batches([v[0].shape[0] for v in videos], self.batch_size, self.ramp_start)
        x, s, f = self._pre_out
        for bi, segs in enumerate(batches):
            fu, su, fl, ev_in, ev_free = self._raw[bi % 2]
            with torch.cuda.stream(self._copy_stream):
                if bi >= 2:
                    self._copy_stream.wait_event(ev_free)
                o = 0
                for (vi, b0, n) in segs:
                    frames_u8, segmaps_u8, flow_raw = videos[vi]
                    for dst, src in ((fu, frames_u8), (su, segmaps_u8)) + (((fl, flow_raw),) if fl is not None else ()):
                        oo = o
                        for piece in _pieces(src, b0, n):
                            dst[oo:oo + piece.shape[0]].copy_(piece, non_blocking=True)
                            oo += piece.shape[0]
                    o += n
                self.h2d_bytes += 2 * o * H * W * 3 + (0 if fl is None else o * fhw[0] * fhw[1] * 2 * 4)
                ev_in.record(self._copy_stream)
            compute.wait_event(ev_in)
            self._prep.images(fu[:o], out=x[:o])
            self._prep.images(su[:o], out=s[:o])
            if fl is not None:
                self._prep.flow(fl[:o], out=f[:o])
            ev_free.record(compute)  # the raw staging buffers are free once the transforms have run
            feats = self.model(x[:o], s[:o], None if fl is None else f[:o], return_features=True)
            o = 0
            for (vi, b0, n) in segs:
                outs[vi][b0:b0 + n].copy_(feats[o:o + n], non_blocking=True)
                if device_outs is not None:
                    device_outs[vi][b0:b0 + n].copy_(feats[o:o + n], non_blocking=True)
                o += n
            self.d2h_bytes += o * D * 4
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

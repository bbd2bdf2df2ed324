"""world_size-2 gloo test (CPU) of the N>1 host path of the 80-video job: LPT sharding, frames of a rank's videos packed into batches
that cross video boundaries, every rank writing its [T_v, D] blocks straight into the rows of ONE shared-memory LFB array
(lfb.SharedLFB) — the host-side gather in video order with no collective on the data path (generate_evp_LFB.py:457)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import surgvid_b200  # noqa: F401
from surgvid_b200 import lfb

LENGTHS = [37, 12, 55, 20, 41, 9, 30]
D = 16


def _frames(video):
    g = torch.Generator().manual_seed(2000 + video)
    return torch.randn(LENGTHS[video], 3, generator=g)


def _fake_model(x):
    """Deterministic per-frame stand-in of the encoder ([n, 3] -> [n, D]); the CUDA forward is covered by the -m gpu tests."""
    w = torch.linspace(-1.0, 1.0, 3 * D).view(3, D)
    return torch.tanh(x @ w)


def _worker(rank, world, port, name):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assign = lfb.lpt_assign(LENGTHS, world)
    shared = lfb.SharedLFB(name, LENGTHS, D, create=(rank == 0)) if rank == 0 else None
    dist.barrier()
    if rank != 0:
        shared = lfb.SharedLFB(name, LENGTHS, D)
    mine = assign[rank]
    vids = [_frames(v) for v in mine]
    outs = shared.blocks(mine)
    for segs in lfb.pack_batches([LENGTHS[v] for v in mine], 16, ramp_start=4):
        xb = torch.cat([vids[vi][b0:b0 + n] for (vi, b0, n) in segs])
        fb = _fake_model(xb)
        o = 0
        for (vi, b0, n) in segs:
            outs[vi][b0:b0 + n].copy_(fb[o:o + n])
            o += n
    dist.barrier()                                   # bookkeeping only: every rank's blocks are in place
    if rank == 0:
        expect = torch.cat([_fake_model(_frames(v)) for v in range(len(LENGTHS))])
        assert torch.equal(shared.array, expect)    # bit-identical to the single-process order
    dist.barrier()
    shared.unlink() if rank == 0 else shared.close()
    dist.destroy_process_group()


def test_two_rank_shared_lfb_gather_matches_single_process():
    ctx = mp.get_context("spawn")
    port = 29700 + (os.getpid() % 2000)
    name = f"surgvid_cpu_test_{os.getpid()}"
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0


def test_pack_batches_covers_every_frame_once_and_crosses_videos():
    for lengths, b, r in (([37, 12, 55], 16, 4), ([5], 8, None), ([800, 800, 700], 800, 100), ([1, 1, 1], 2, None)):
        batches = lfb.pack_batches(lengths, b, r)
        seen = [np.zeros(n, dtype=np.int64) for n in lengths]
        for segs in batches:
            assert 0 < sum(s[2] for s in segs) <= b
            for (vi, b0, n) in segs:
                seen[vi][b0:b0 + n] += 1
        assert all((s == 1).all() for s in seen)
        if r is None:   # no ramp: every batch but the last is full
            assert all(sum(s[2] for s in segs) == b for segs in batches[:-1])
    assert any(len(segs) > 1 for segs in lfb.pack_batches([37, 12, 55], 16))
    assert lfb.pack_batches([2300], 800, 100) == [[(0, a, n)] for a, n in lfb.ramp_schedule(2300, 800, 100)]


def test_cyclic_frames_pieces():
    pool = torch.arange(10).view(10, 1)
    c = lfb.CyclicFrames(pool, 7, 13)
    assert c.shape == (13, 1)
    assert torch.cat(list(c.pieces(0, 13))).flatten().tolist() == [(7 + t) % 10 for t in range(13)]
    assert torch.cat(list(c.pieces(2, 3))).flatten().tolist() == [9, 0, 1]

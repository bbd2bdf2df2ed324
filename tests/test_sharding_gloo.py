"""world_size-2 gloo test of the N>1 path's host logic: videos are LPT-sharded across ranks with NO data-path collective;
each rank extracts its own videos; rank 0 gathers the per-video blocks in video order.  The per-frame 'model' here is a
deterministic stand-in so the test runs on CPU; the CUDA forward is covered by the -m gpu tests."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import surgvid_b200  # noqa: F401
from surgvid_b200 import lfb


def _fake_features(video: int, T: int) -> np.ndarray:
    rng = np.random.default_rng(1000 + video)
    return rng.standard_normal((T, 16)).astype(np.float32)


def _worker(rank, world, port, lengths, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assign = lfb.lpt_assign(lengths, world)
    mine = [_fake_features(v, lengths[v]) for v in assign[rank]]
    frames = torch.tensor([sum(lengths[v] for v in assign[rank])], dtype=torch.int64)
    dist.barrier()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)      # host-side gather of feature blocks (not on the timed data path)
    total = frames.clone()
    dist.all_reduce(total)                          # bookkeeping only: frames processed by all ranks
    if rank == 0:
        out = lfb.gather_in_video_order(gathered, assign, len(lengths))
        q.put((out, int(total)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_matches_single_process():
    lengths = [37, 12, 55, 20, 41, 9, 30]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.concatenate([_fake_features(v, lengths[v]) for v in range(len(lengths))])
    assert total == sum(lengths)
    assert np.array_equal(out, expect)  # bit-identical to the single-process order

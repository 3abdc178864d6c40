"""CPU, world_size 2 over gloo: frame sharding and the final score gather (the only communication of the
inference path).  No kernels are launched."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cwfa_b200.sharding import frame_shard, gather_frame_scores


def test_shards_partition_frames():
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 4, 8):
            blocks = [frame_shard(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert frame_shard(1024, 3, 8) == (384, 512)          # 128 frames per GPU (BASELINE configs[4])


def _worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = frame_shard(n_frames, rank, world)
    # stand-in for per-frame (NLL, logdet): a deterministic function of the frame id
    local = torch.stack([torch.tensor([float(f), float(f) ** 2]) for f in range(a, b)]) if b > a else torch.zeros(0, 2)
    full = gather_frame_scores(local, n_frames)
    t = torch.tensor([float(b - a)])
    dist.all_reduce(t)                                    # same reduction bench.py uses for frames processed
    if rank == 0:
        q.put((full, float(t)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [5, 8])
def test_two_rank_gather_matches_single_process(n_frames):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    [p.start() for p in procs]
    full, total = q.get()
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    ref = torch.stack([torch.tensor([float(f), float(f) ** 2]) for f in range(n_frames)])
    assert torch.equal(full, ref) and total == n_frames

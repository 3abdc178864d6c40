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
        q.put((full.tolist(), float(t)))          # by value: a tensor would travel as a shared-memory fd that dies with this process
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
    assert torch.equal(torch.tensor(full).reshape(ref.shape), ref) and total == n_frames


# ---- training: gradient all-reduce of the flat buffers (cwfa_b200/training.py) -----------------------
def _toy_net():
    torch.manual_seed(7)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 6, 3, padding=1), torch.nn.PReLU(), torch.nn.Conv2d(6, 2, 1))


def _toy_grads(net, frame_id):
    g = torch.Generator().manual_seed(100 + frame_id)
    x = torch.randn(1, 3, 8, 8, generator=g)
    for p in net.parameters():
        if p.grad is not None:
            p.grad.zero_()
    net(x).square().mean().backward()


def _train_worker(rank, world, port, q, overlap=False):
    from cwfa_b200.training import Lion, allreduce_gradients, allreduce_nll_terms
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net = _toy_net()
    opt = Lion([{"params": [net[0].weight, net[0].bias]}, {"params": list(net[1].parameters()) + list(net[2].parameters())}], lr=1e-3)
    if overlap:
        opt.enable_overlap(bucket_bytes=64)                 # tiny buckets: several collectives per flat buffer, launched from the hooks
    if overlap and rank == 0:                               # a backward outside an armed step (one rank only) issues no collective
        opt.zero_grad()
        _toy_grads(net, 5)
    opt.zero_grad()
    if overlap:
        opt.arm_overlap()                                   # what a trainer's step() does before its backward
    _toy_grads(net, rank)                                   # every rank its own frame
    n = allreduce_gradients([opt])
    tot = allreduce_nll_terms(torch.tensor([1.0 + rank, 2.0]), torch.tensor([10.0 * (rank + 1), 1.0]))
    if rank == 0:
        q.put((n, opt.grad_scale, [g.tolist() for g in opt.flat_grads()], [float(t) for t in tot]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [False, True])
def test_two_rank_gradient_allreduce_matches_single_process(overlap):
    """``overlap``: the bucketed all-reduce launched from the post-accumulate hooks DURING backward (FlatGroup.enable_overlap)
    must give the same flat gradients as one collective after backward."""
    from cwfa_b200.training import Lion
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q, overlap)) for r in range(2)]
    [p.start() for p in procs]
    n, scale, flats, tot = q.get()
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert (n > 2 if overlap else n == 2) and scale == 0.5  # one collective per flat buffer (or per bucket); the mean is folded into Lion
    # single process: sum of the two frames' gradients in the same flat layout
    net = _toy_net()
    opt = Lion([{"params": [net[0].weight, net[0].bias]}, {"params": list(net[1].parameters()) + list(net[2].parameters())}], lr=1e-3)
    acc = [torch.zeros_like(g) for g in opt.flat_grads()]
    for f in range(2):
        opt.zero_grad()
        _toy_grads(net, f)
        for a, g in zip(acc, opt.flat_grads()):
            a += g
    for a, g in zip(acc, flats):
        g = torch.tensor(g)
        assert torch.allclose(a, g, rtol=1e-6, atol=1e-8) and float(g.abs().sum()) > 0
    assert tot == [3.0 + 4.0, 30.0 + 2.0, 4.0]

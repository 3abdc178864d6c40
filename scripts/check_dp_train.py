"""Data-parallel training check (BASELINE.json configs[3]: "grad parity vs single-GPU"), run under torchrun with N >= 2 ranks:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/check_dp_train.py

Every rank trains flow level 0 of a small model on ITS OWN frame (frames sharded 1/rank) with the flat gradient buffers
all-reduced over NCCL.  Rank 0 also runs the same K steps in a single process (gradients of all N frames accumulated, mean
folded into Lion) from the same initial weights, and the two parameter sets are compared.
"""
import copy
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cwfa_b200                                                        # noqa: E402
from cwfa_b200.training import FlowLevelTrainer, flow_level_loss       # noqa: E402

K, D, S = 3, 16, 64
GRAPH = "--graph" in sys.argv          # graph=True on the trainers: falls back to the eager step under data parallelism


def frame(f, dev):
    g = torch.Generator().manual_seed(500 + f)
    mk = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)
    return mk(1, D, S, S), mk(1, 29, S, S), mk(1, D // 2, S, S, sc=0.1), mk(1, D // 2, S, S)


def params_of(model):
    return torch.cat([p.detach().reshape(-1).float().cpu() for p in list(model.conv_inn[0].parameters()) + list(model.cond_nets[0].parameters())
                      if p.requires_grad])


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    model = cwfa_b200.CWFAModel(n_depths=D, volume_side_size=S, INN_max_down_steps=2, seed=0).to(dev)
    init = copy.deepcopy(model.state_dict())
    ref = None
    if rank == 0:                                                      # single-process reference, same initial weights
        tr = FlowLevelTrainer(model, 0, lr=1e-4, lr_cond=1e-4)
        for _ in range(K):
            tr.optimizer.zero_grad(); tr.optimizer_cond.zero_grad()
            for f in range(world):
                flow_level_loss(model, 0, *frame(f, dev))[0].backward()
            tr.optimizer.grad_scale = tr.optimizer_cond.grad_scale = 1.0 / world
            tr.optimizer_cond.step(); tr.optimizer.step()
        ref = params_of(model)
        tr.release()
        model.load_state_dict(init)
    dist.barrier()
    tr = FlowLevelTrainer(model, 0, lr=1e-4, lr_cond=1e-4, graph=GRAPH)
    inputs = frame(rank, dev)
    losses = [float(tr.step(*inputs)["loss"]) for _ in range(K)]
    mine = params_of(model)
    gathered = [torch.empty_like(mine).to(dev) for _ in range(world)]
    dist.all_gather(gathered, mine.to(dev))
    if rank == 0:
        same_on_all_ranks = all(torch.equal(gathered[0], g) for g in gathered)
        out = {"world": world, "steps": K, "graph": GRAPH, "collectives_per_step": tr.collectives, "replicas_identical": bool(same_on_all_ranks),
               "max_abs_param_diff_vs_single_process": float((mine - ref).abs().max()),
               "params_changed_by_training": float((ref - params_of_init(init, model)).abs().max()), "losses_rank0": losses}
        print(json.dumps(out))
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/dp_train_check.json", "w") as f:
            json.dump(out, f, indent=1)
        assert same_on_all_ranks and out["max_abs_param_diff_vs_single_process"] <= 2e-4 * 1.0     # <= two Lion steps of lr 1e-4
    dist.barrier()
    dist.destroy_process_group()


def params_of_init(init, model):
    names = [("conv_inn.0." + k) for k, p in model.conv_inn[0].named_parameters() if p.requires_grad] + \
            [("cond_nets.0." + k) for k, p in model.cond_nets[0].named_parameters() if p.requires_grad]
    return torch.cat([init[k].reshape(-1).float().cpu() for k in names])


if __name__ == "__main__":
    main()

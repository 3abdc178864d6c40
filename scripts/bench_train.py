"""Training step of ONE flow level at the full config (BASELINE.json configs[3], fp32 module path): forward NLL + inverse
MSE + backward + Lion, timed with CUDA events.  Under torchrun every rank trains on its own frame and the flat gradient
buffers are all-reduced over NCCL (frames sharded 1/rank).

    python scripts/bench_train.py --level 0 --steps 5 --warmup 2 [--side 512] [--json out.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cwfa_b200                                            # noqa: E402
from cwfa_b200 import _lib                                  # noqa: E402
from cwfa_b200.training import FlowLevelTrainer             # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=0)
    ap.add_argument("--lrnn", action="store_true", help="time the LRNN ('last step') training step instead of a flow level")
    ap.add_argument("--side", type=int, default=512)
    ap.add_argument("--depths", type=int, default=96)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16", "fp16"])
    ap.add_argument("--json", default=None)
    ap.add_argument("--graph", action="store_true", help="capture the whole step as a CUDA graph (trainer graph=True)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(dev))
    n, S, D = a.level, a.side, a.depths
    steps_total = 5 if a.lrnn else n + 2
    model = cwfa_b200.CWFAModel(n_depths=D, volume_side_size=S, INN_max_down_steps=steps_total, seed=0).to(dev)
    if a.lrnn:
        from cwfa_b200.training import LRNNTrainer
        lt = LRNNTrainer(model, precision=a.precision, graph=a.graph)
    else:
        tr = FlowLevelTrainer(model, n, precision=a.precision, graph=a.graph)
    C = D // 2 ** n
    g = torch.Generator(device="cpu").manual_seed(1000 + rank)             # every rank its own frame
    mk = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)
    gt, views = mk(a.batch, C, S, S), mk(a.batch, 29, S, S)
    mean_vol, vol_in = mk(a.batch, C // 2, S, S, sc=0.1), mk(a.batch, C // 2, S, S)
    if a.lrnn:
        gt_l = mk(a.batch, D // 16, S, S)

        class _T:                                              # same call shape as FlowLevelTrainer for the loops below
            collectives = 0
            def step(self, *_):
                r = lt.step(gt_l, views)
                self.collectives = lt.collectives
                return r
        tr = _T()
    losses = []
    for _ in range(a.warmup):
        losses.append(float(tr.step(gt, views, mean_vol, vol_in)["loss"]))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        parts = tr.step(gt, views, mean_vol, vol_in)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    losses.append(float(parts["loss"]))
    if rank == 0:
        out = {"metric": "flow-level training steps/s (fwd NLL + inverse MSE + backward + Lion)", "precision": a.precision, "graph": bool(a.graph), "level": "lrnn" if a.lrnn else n,
               "value": world * a.batch * 1000.0 / ms, "unit": "frames/s", "ms_per_step": ms, "n_gpus": world, "batch_per_gpu": a.batch,
               "side": S, "depths": D, "launches_per_step": (_lib.launch_count - l0) / a.steps, "collectives_per_step": tr.collectives,
               "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "losses": losses}
        print(json.dumps(out))
        if a.json:
            with open(a.json, "w") as f:
                json.dump(out, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

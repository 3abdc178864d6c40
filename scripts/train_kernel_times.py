"""Warm per-entry-point GPU time of ONE flow-level training step (bf16): every C-ABI call bracketed by CUDA events on its stream
(eager; the gaps between calls -- torch element-wise ops, host launch latency -- are reported as the remainder)."""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cwfa_b200
from cwfa_b200 import _lib
from cwfa_b200.training import FlowLevelTrainer, LRNNTrainer

dev = "cuda:0"
S, D = 512, 96
lrnn = "--lrnn" in sys.argv
model = cwfa_b200.CWFAModel(n_depths=D, volume_side_size=S, INN_max_down_steps=5 if lrnn else 2, seed=0).to(dev)
g = torch.Generator().manual_seed(1)
mk = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)
if lrnn:
    tr = LRNNTrainer(model, precision="bf16")
    args = (mk(1, D // 16, S, S), mk(1, 29, S, S))
else:
    tr = FlowLevelTrainer(model, 0, precision="bf16")
    args = (mk(1, D, S, S), mk(1, 29, S, S), mk(1, D // 2, S, S, sc=0.1), mk(1, D // 2, S, S))
for _ in range(3):
    tr.step(*args)
torch.cuda.synchronize()
evs = []
orig = _lib.call


def timed(name, *a):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    r = orig(name, *a)
    e.record()
    evs.append((name, s, e))
    return r


_lib.call = timed
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
tr.step(*args)
e1.record()
torch.cuda.synchronize()
_lib.call = orig
agg = collections.defaultdict(lambda: [0, 0.0])
for name, s, e in evs:
    agg[name][0] += 1
    agg[name][1] += s.elapsed_time(e)
tot = sum(v for _, v in agg.values())
print(f"step {e0.elapsed_time(e1):.2f} ms (with event overhead); C-ABI calls {len(evs)}, {tot:.2f} ms inside them")
for name, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
    print(f"{v:8.3f} ms  x{c:4d}  {name}")

"""Per-call time of every C-ABI call in one eager forward_nll pass at batch B (CUDA events around each call)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cwfa_b200
from cwfa_b200 import _lib, tc, ops
from cwfa_b200.engine import CWFAEngine
from bench import synthetic_inputs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = dict(side=512, depths=96, steps=5)
dev = torch.device("cuda:0")
model = cwfa_b200.CWFAModel(n_depths=96, volume_side_size=512, INN_max_down_steps=5, seed=0).to(dev)
eng = CWFAEngine(model, "bf16")
views, mvs = synthetic_inputs(cfg, dev, 100)
vol = torch.randn(B, 96, 512, 512, device=dev)
vB = views.to(dev).repeat(B, 1, 1, 1)
mvB = [m.to(dev).repeat(B, 1, 1, 1) for m in mvs[:model.n_levels]]
for _ in range(2):
    eng.forward_nll(vol, vB, mvB)
torch.cuda.synchronize()
evs = []
orig = _lib.call
def timed(name, *a):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); orig(name, *a); e.record()
    evs.append((name, s, e))
for mod in (_lib, tc, ops):
    if hasattr(mod, "_lib"): mod._lib.call = timed
_lib.call = timed
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.forward_nll(vol, vB, mvB); e1.record()
torch.cuda.synchronize()
agg = collections.OrderedDict()
for k, s, e in evs:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += s.elapsed_time(e)
tot = sum(a[1] for a in agg.values())
print(f"forward_nll B={B}: wall {e0.elapsed_time(e1):.1f} ms; sum of calls {tot:.1f} ms")
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:9.2f} ms {100*t/tot:5.1f}% x{c:3d}  {k}")

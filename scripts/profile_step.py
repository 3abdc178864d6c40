"""One eager reconstruct step bracketed by cudaProfilerStart/Stop (use: ncu --profile-from-start off ...)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cwfa_b200
from cwfa_b200.engine import CWFAEngine
from bench import synthetic_inputs
side = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfg = dict(side=side, depths=96, steps=5)
dev = torch.device("cuda:0")
model = cwfa_b200.CWFAModel(n_depths=96, volume_side_size=side, INN_max_down_steps=5, seed=0).to(dev)
eng = CWFAEngine(model, "bf16")
views, mvs = synthetic_inputs(cfg, dev, 100)
views, mvs = views.to(dev), [m.to(dev) for m in mvs]
for _ in range(2):
    eng.reconstruct(views, mvs)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
eng.reconstruct(views, mvs)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")

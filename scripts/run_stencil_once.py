"""A few launches of the fused depth-stencil kernel at the level-0 shape (512 x 512 x 48) for `ncu -k regex:stencil3d`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cwfa_b200 import tc

DEV = "cuda:0"
D = int(sys.argv[1]) if len(sys.argv) > 1 else 48
w1, b1 = torch.randn(32, 1, 3, 3, 3, device=DEV) * 0.3, torch.randn(32, device=DEV) * 0.1
w2, b2 = torch.randn(1, 32, 3, 3, 3, device=DEV) * 0.1, torch.randn(1, device=DEV)
sl = torch.tensor([0.25], device=DEV)
sw = tc.StencilWeights(w1, b1, w2, b2, "bf16")
xs = [tc.to_c8(torch.randn(1, D, 512, 512, device=DEV)) for _ in range(3)]
for i in range(3):
    tc.stencil3d_tc(xs[i], sw, sl, D)
torch.cuda.synchronize()
print("ok")

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cwfa_b200 import tc
DEV = "cuda:0"
x = tc.to_c8(torch.randn(1, 64, 512, 512, device=DEV))
p3 = tc.PackedConv(torch.randn(64, 64, 3, 3, device=DEV) * 0.04, torch.zeros(64, device=DEV), bn=64)
p1 = tc.PackedConv(torch.randn(64, 64, 1, 1, device=DEV) * 0.1, torch.zeros(64, device=DEV), bn=64)
for _ in range(4): y = tc.resblock_tc(x, p3, p1)
torch.cuda.synchronize(); print("ok")

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cwfa_b200 import tc
DEV = "cuda:0"
ch = 48
b = tc.to_c8(torch.randn(1, 64, 512, 512, device=DEV))
pc = tc.PackedConv(torch.randn(2 * ch, 64, 3, 3, device=DEV) * 0.04, torch.zeros(2 * ch, device=DEV), bn=96)
x = torch.randn(1, ch, 512, 512, device=DEV)
perm = torch.randperm(ch, device=DEV).to(torch.int32)
ld = torch.zeros(1, device=DEV)
for _ in range(4):
    y = tc.conv_tc_coupling(b, pc, x, ch=ch, inverse=True, perm=perm, perm_axis=1, logdet=ld)
torch.cuda.synchronize(); print("ok")

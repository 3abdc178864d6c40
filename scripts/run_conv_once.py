"""One conv_tc shape a few times (for ncu): run_conv_once.py cin cout k mb bn [act]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cwfa_b200 import tc, ops
DEV = "cuda:0"
cin, cout, k, mb, bn = [int(a) for a in sys.argv[1:6]]
act = int(sys.argv[6]) if len(sys.argv) > 6 else ops.ACT_PRELU
x = tc.to_c8(torch.randn(1, cin, 512, 512, device=DEV))
pc = tc.PackedConv(torch.randn(cout, cin, k, k, device=DEV) * (1.0 / (cin * k * k)) ** 0.5, torch.zeros(cout, device=DEV), bn=bn)
slope = torch.full((1,), 0.25, device=DEV)
for _ in range(3):
    y = tc.conv_tc(x, pc, act=act, slope=slope, mb=mb)
torch.cuda.synchronize(); print("ok")

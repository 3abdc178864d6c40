import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cwfa_b200 import tc, _lib
DEV = "cuda:0"
H = W = 512
x = tc.to_c8(torch.randn(1, 64, H, W, device=DEV))
p3 = tc.PackedConv(torch.randn(64, 64, 3, 3, device=DEV) * 0.04, torch.zeros(64, device=DEV), bn=64)
p1 = tc.PackedConv(torch.randn(64, 64, 1, 1, device=DEV) * 0.1, torch.zeros(64, device=DEV), bn=64)
for _ in range(3): tc.resblock_tc(x, p3, p1)
dbg = torch.zeros(148 * 8 * 8, dtype=torch.int64, device=DEV)
_lib.call("cwfa_resblock_set_debug_buffer", dbg.data_ptr())
tc.resblock_tc(x, p3, p1); torch.cuda.synchronize()
_lib.call("cwfa_resblock_set_debug_buffer", None)
d = dbg.view(148, 8, 8).cpu().double()
t0 = d[:, 0, 0].min()
names = ["g1_issue_start", "g1_issue_end", "epi1 sees acc1_full", "epi1 a2_empty ok", "epi1 mb0 ld done", "epi1 done", "epi2 sees acc2_full", "epi2 done"]
for cta in (0, 77):
    print("CTA", cta)
    for i in range(7):
        print("  tile", i, " ".join(f"{(d[cta, i, k] - t0) / 1e3:7.2f}" for k in (0,1,2,3,7,4,5,6)))
print("columns:", names)

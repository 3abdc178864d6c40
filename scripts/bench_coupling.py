"""Graph-replay timing of the persistent fused coupling kernel at the frame's shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cwfa_b200 import tc
DEV = "cuda:0"
for ch, axis in ((48, 1), (48, 3), (48, 0), (24, 1), (12, 2), (6, 1)):
    bs = [tc.to_c8(torch.randn(1, 64, 512, 512, device=DEV)) for _ in range(6)]
    pc = tc.PackedConv(torch.randn(2 * ch, 64, 3, 3, device=DEV) * 0.04, torch.zeros(2 * ch, device=DEV), bn=tc.pad16(2 * ch))
    xs = [torch.randn(1, ch, 512, 512, device=DEV) for _ in range(6)]
    n_ax = {0: 1, 1: ch, 2: 512, 3: 512}[axis]
    perm = torch.randperm(n_ax, device=DEV).to(torch.int32) if axis else None
    ld = torch.zeros(1, device=DEV)
    fn = lambda i: tc.conv_tc_coupling(bs[i], pc, xs[i], ch=ch, inverse=True, perm=perm, perm_axis=axis, logdet=ld)
    for i in range(6): fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = [fn(i) for i in range(6)]
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 6)
    print(f"coupling ch={ch} perm axis {axis}: {best*1e3:.1f} us (incl. finalize)", flush=True)
    del keep, g

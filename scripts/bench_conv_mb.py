"""U-Net conv shapes with MB = 1 (two CTAs per SM, 256 TMEM columns each) against MB = 2 (one CTA, 512 columns, epilogue exposed),
with and without the BatchNorm statistics in the epilogue; CUDA events around a graph replay, inputs rotated through > L2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cwfa_b200 import tc, ops
DEV = "cuda:0"

def bench(fn, n):
    for i in range(n):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = [fn(i) for i in range(n)]
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best * 1e3

sl = torch.tensor([0.2], device=DEV)
for c, s in ((256, 512), (512, 256), (1024, 128)):
    n = 4
    xs = [tc.to_c8(torch.randn(1, c, s, s, device=DEV)) for _ in range(n)]
    pc = tc.PackedConv(torch.randn(c, c, 3, 3, device=DEV) * 0.02, torch.zeros(c, device=DEV))
    fl = 2.0 * s * s * c * c * 9
    out = [f"{c}->{c} 3x3 @{s}: BN={pc.BN}"]
    for mb in (2, 1):
        us = bench(lambda i: tc.conv_tc(xs[i], pc, act=ops.ACT_PRELU, slope=sl, mb=mb), n)
        out.append(f"mb={mb}: {us:.1f} us ({fl / us / 1e6:.0f} TF/s)")
    us = bench(lambda i: tc.conv_tc_bn_stats(xs[i], pc, act=ops.ACT_PRELU, slope=sl, mb=2)[0], n)
    out.append(f"mb=2 + stats: {us:.1f} us")
    print("   ".join(out), flush=True)

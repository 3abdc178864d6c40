"""One wgrad_tc launch at the full-config trunk shape (64 -> 64, 3x3, 512x512) for `ncu --set full`; with --time it prints
CUDA-event timings (not under the profiler) of the trunk / final / banded-stencil shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cwfa_b200 import tc, autograd as ag
DEV = "cuda:0"
shapes = [(64, 64, 3), (64, 96, 3), (64, 64, 1), (48, 1536, 3), (1536, 48, 3)] if "--time" in sys.argv else [(64, 64, 3)]
for Cin, Cout, K in shapes:
    x = tc.to_c8(torch.randn(1, Cin, 512, 512, device=DEV))
    dy = tc.to_c8(torch.randn(1, Cout, 512, 512, device=DEV))
    for _ in range(3):
        dw = ag.conv2d_wgrad_tc(x, dy, Cin, Cout, K)
    if "--time" in sys.argv:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
        ts = []
        for _ in range(5):
            flush.zero_()
            e0.record(); dw = ag.conv2d_wgrad_tc(x, dy, Cin, Cout, K); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
        us = sorted(ts)[len(ts) // 2]
        fl = 2.0 * 512 * 512 * Cin * Cout * K * K
        by = 2.0 * 512 * 512 * (tc.pad16(Cin) + tc.pad16(Cout))
        print(f"wgrad_tc {Cin}->{Cout} {K}x{K} @512x512: {us:.1f} us (kernel + finalize, L2 flushed)  {fl / us / 1e6:.1f} TFLOP/s  "
              f"operand bytes {by / 1e6:.0f} MB -> {by / us / 1e3:.0f} GB/s")
torch.cuda.synchronize(); print("ok")

"""Graph-replay timing of the transposed convolutions of the U-Net up path (with / without the skip add)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cwfa_b200 import tc
DEV = "cuda:0"
for cin, cout, H in ((1024, 512, 128), (512, 256, 256)):
    n = 6
    xs = [tc.to_c8(torch.randn(1, cin, H, H, device=DEV)) for _ in range(n)]
    skips = [tc.to_c8(torch.randn(1, cout, 2 * H, 2 * H, device=DEV)) for _ in range(n)]
    w = torch.randn(cin, cout, 2, 2, device=DEV) * (1.0 / cin) ** 0.5
    for bn in (256, 128):
        pc = tc.PackedConv(w, torch.zeros(cout, device=DEV), transposed=True, bn=bn)
        for mb in (2, 1):
            for with_skip in (True, False):
                fn = lambda i: tc.conv_transpose_tc(xs[i], pc, skips[i] if with_skip else None, mb=mb)
                for i in range(n): fn(i)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    keep = [fn(i) for i in range(n)]
                g.replay(); torch.cuda.synchronize()
                best = 1e9
                for _ in range(4):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1) / n)
                fl = 2.0 * H * H * cin * cout * 4
                print(f"convT {cin}->{cout} @{H}x{H} bn{bn} mb{mb} skip={with_skip}: {best*1e3:.1f} us  {fl/best/1e9:.0f} TFLOP/s", flush=True)
                del keep, g

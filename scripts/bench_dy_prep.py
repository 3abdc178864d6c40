"""dy_prep (fused ELU adjoint + C8 conversion + bias gradient) against the three passes it replaces, 64 ch x 512 x 512."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cwfa_b200 import _lib, ops, tc
from cwfa_b200 import autograd as ag

DEV = "cuda:0"


def timed(fn, n):
    for i in range(n):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = [fn(i) for i in range(n)]
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    del keep, g
    return best * 1e3


n = 6
for C in (64, 96):
    dys = [torch.randn(1, C, 512, 512, device=DEV) for _ in range(n)]
    ys = [torch.randn(1, C, 512, 512, device=DEV) for _ in range(n)]
    def old(i):
        g = torch.empty_like(dys[i])
        _lib.call("cwfa_elu_bwd_f32", dys[i].data_ptr(), ys[i].data_ptr(), g.data_ptr(), g.numel(), ag._stream())
        return tc.to_c8(g, "bf16"), ag.channel_sum(g)
    print(f"C={C}: elu_bwd + nchw_to_c8 + channel_sum: {timed(old, n):.1f} us")
    print(f"C={C}: dy_prep (ELU, C8, bias):            {timed(lambda i: tc.dy_prep(dys[i], ys[i], 'bf16', want_bias=True), n):.1f} us")
    print(f"C={C}: dy_prep (ELU, C8, bias, +fp32 out): {timed(lambda i: tc.dy_prep(dys[i], ys[i], 'bf16', want_bias=True, want_f32=True), n):.1f} us")
    print(f"C={C}: dy_prep (C8 only):                  {timed(lambda i: tc.dy_prep(dys[i], None, 'bf16'), n):.1f} us")
    print(f"C={C}: nchw_to_c8 alone:                   {timed(lambda i: tc.to_c8(dys[i], 'bf16'), n):.1f} us")
print("ok")

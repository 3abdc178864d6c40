"""Timing of the fused depth-stencil kernel at the full-config shapes (512 x 512, D = 48 / 24 / 12 / 6), CUDA events around a CUDA-graph
replay on rotated inputs, against the banded two-convolution form.   python scripts/bench_stencil3d.py [rows_max ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cwfa_b200 import tc

DEV = "cuda:0"
rows_list = [int(a) for a in sys.argv[1:]] or [0]


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
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    del keep, g
    return best * 1e3


for D in (48, 24, 12, 6):
    n = 6
    w1, b1 = torch.randn(32, 1, 3, 3, 3, device=DEV) * 0.3, torch.randn(32, device=DEV) * 0.1
    w2, b2 = torch.randn(1, 32, 3, 3, 3, device=DEV) * 0.1, torch.randn(1, device=DEV)
    sl = torch.tensor([0.25], device=DEV)
    sw = tc.StencilWeights(w1, b1, w2, b2, "bf16")
    xs = [tc.to_c8(torch.randn(1, D, 512, 512, device=DEV)) for _ in range(n)]
    for rm in rows_list:
        us = timed(lambda i: tc.stencil3d_tc(xs[i], sw, sl, D, rows_max=rm), n)
        vox = 512 * 512 * D
        print(f"stencil3d_tc D={D} rows_max={rm or 768}: {us:.1f} us  ({vox / us / 1e3:.1f} Gvoxel/s, {vox * 3456 / us / 1e6:.1f} TFLOP/s true 3-D flops)", flush=True)
    # the banded two-convolution form it replaces (packed._CondNet with fuse_stencil off)
    from cwfa_b200 import ops
    Cm = 32
    W1 = torch.zeros(D, Cm, D, 3, 3, device=DEV)
    W2 = torch.zeros(D, D, Cm, 3, 3, device=DEV)
    for kd in range(3):
        for d in range(D):
            dp = d + kd - 1
            if 0 <= dp < D:
                W1[d, :, dp] = w1[:, 0, :, :, kd]
                W2[d, dp] = w2[0, :, :, :, kd]
    s1 = tc.PackedConv(W1.reshape(D * Cm, D, 3, 3), b1.repeat(D), "bf16")
    Wg = tc.col2im3x3_weights(W2.reshape(D, D * Cm, 3, 3))
    gp = tc.pad16(Wg.shape[0])
    s2g = tc.PackedConv(Wg, None, "bf16", bn=144 if gp % 144 == 0 else gp)
    bias = torch.zeros(tc.pad16(D), device=DEV)
    mb = 1 if s2g.BN == 144 else 2
    us = timed(lambda i: tc.col2im3x3_c8(tc.conv_tc(tc.conv_tc(xs[i], s1, act=ops.ACT_PRELU, slope=sl), s2g, mb=mb), bias, D), n)
    print(f"banded form   D={D}: {us:.1f} us (s1 conv + s2 1x1 conv + col2im)", flush=True)
    del s1, s2g, W1, W2, Wg
print("ok")
# phase breakdown of CTA (0, 0) (cycles per pixel-row), level 0
from cwfa_b200 import _lib
buf = torch.zeros(8, device=DEV, dtype=torch.int64)
_lib.call("cwfa_stencil_set_debug_buffer", buf.data_ptr())
tc.stencil3d_tc(xs[0] if D == 48 else tc.to_c8(torch.randn(1, 48, 512, 512, device=DEV)), tc.StencilWeights(w1, b1, w2, b2, "bf16"), sl, 48)
torch.cuda.synchronize()
_lib.call("cwfa_stencil_set_debug_buffer", None)
b = buf.tolist()
names = ["-", "im2col", "GEMM1", "epilogue 1", "GEMM2", "epilogue 2", "gather"]
print("phase cycles per pixel-row (CTA 0, D=48):", {n: round(v / max(b[7], 1)) for n, v in zip(names, b[:7])}, "pixel-rows", b[7], "total/row", round(sum(b[:7]) / max(b[7], 1)))

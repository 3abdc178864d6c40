"""A few launches of ONE hot-path kernel at its full-config shape, for `ncu --set full -k regex:<name>`; with --time it
prints CUDA-event timings (graph replay, inputs rotated through > L2; never under the profiler).

    python scripts/run_kernel_once.py coupling|resblock|stencil|bn [--time]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cwfa_b200 import ops, tc

DEV = "cuda:0"
what = sys.argv[1] if len(sys.argv) > 1 else "coupling"
timed = "--time" in sys.argv
P = 512 * 512


def bench(fn, n):
    for i in range(n):
        fn(i)
    torch.cuda.synchronize()
    if not timed:
        return None
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


if what == "coupling":
    for ch in ((48, 24, 12, 6) if timed else (48,)):
        n = 6
        bs = [tc.to_c8(torch.randn(1, 64, 512, 512, device=DEV)) for _ in range(n)]
        pc = tc.PackedConv(torch.randn(2 * ch, 64, 3, 3, device=DEV) * 0.04, torch.zeros(2 * ch, device=DEV), bn=tc.pad16(2 * ch))
        xs = [torch.randn(1, ch, 512, 512, device=DEV) for _ in range(n)]
        perm = torch.randperm(ch, device=DEV).to(torch.int32)
        ld = torch.zeros(1, device=DEV)
        tk = torch.zeros(8, device=DEV, dtype=torch.int32)
        us = bench(lambda i: tc.conv_tc_coupling(bs[i], pc, xs[i], ch=ch, inverse=True, perm=perm, perm_axis=1, logdet=ld, ticket=tk[:1]), n)
        if us:
            fl = 2.0 * P * 64 * 2 * ch * 9
            print(f"coupling_tc ch={ch} (ticket finalize): {us:.1f} us  {fl / us / 1e6:.0f} TFLOP/s", flush=True)
        us = bench(lambda i: tc.conv_tc_coupling(bs[i], pc, xs[i], ch=ch, inverse=True, perm=perm, perm_axis=1, logdet=ld), n)
        if us:
            print(f"coupling_tc ch={ch} (+ finalize launch): {us:.1f} us", flush=True)
elif what == "coupling_f8":
    for ch in ((48, 24, 12, 6) if timed else (48,)):
        n = 6
        c8 = tc.ch8(ch)
        bs = [tc.to_c8(torch.randn(1, 64, 512, 512, device=DEV)) for _ in range(n)]
        pc = tc.coupling_weights_f8(torch.randn(2 * ch, 64, 3, 3, device=DEV) * 0.04, torch.zeros(2 * ch, device=DEV), ch, torch.randperm(ch), False, "bf16")
        xs = [tc.to_f8(torch.randn(1, ch, 512, 512, device=DEV)) for _ in range(n)]
        perm = torch.randperm(512, device=DEV).to(torch.int32)
        ld = torch.zeros(1, device=DEV)
        tk = torch.zeros(8, device=DEV, dtype=torch.int32)
        us = bench(lambda i: tc.coupling_f8(bs[i], pc, xs[i], ch=ch, inverse=True, perm=perm, perm_axis=2, logdet=ld, ticket=tk[:1]), n)
        if us:
            fl = 2.0 * P * 64 * 2 * ch * 9
            print(f"coupling_f8 ch={ch}: {us:.1f} us  {fl / us / 1e6:.0f} TFLOP/s", flush=True)
elif what == "resblock":
    n = 8
    xin = [tc.to_c8(torch.randn(1, 64, 512, 512, device=DEV)) for _ in range(n)]
    p3 = tc.PackedConv(torch.randn(64, 64, 3, 3, device=DEV) * 0.04, torch.zeros(64, device=DEV), bn=64)
    p1 = tc.PackedConv(torch.randn(64, 64, 1, 1, device=DEV) * 0.1, torch.zeros(64, device=DEV), bn=64)
    us = bench(lambda i: tc.resblock_tc(xin[i], p3, p1), n)
    if us:
        print(f"resblock_tc: {us:.1f} us  {2.0 * P * 64 * 64 * 10 / us / 1e6:.0f} TFLOP/s", flush=True)
elif what == "resblock5":
    n = 4
    xin = [tc.to_c8(torch.randn(1, 320, 512, 512, device=DEV)) for _ in range(n)]
    sets = [(tc.PackedConv(torch.randn(64, 64, 3, 3, device=DEV) * 0.04, torch.zeros(64, device=DEV), bn=64),
             tc.PackedConv(torch.randn(64, 64, 1, 1, device=DEV) * 0.1, torch.zeros(64, device=DEV), bn=64)) for _ in range(5)]
    us = bench(lambda i: tc.resblock_tc_batched(xin[i], sets), n)
    if us:
        print(f"resblock_tc_batched (5 sub-networks): {us:.1f} us = {us / 5:.1f} us per block  {5 * 2.0 * P * 64 * 64 * 10 / us / 1e6:.0f} TFLOP/s", flush=True)
elif what == "convbn":
    n = 4
    xin = [tc.to_c8(torch.randn(1, 256, 512, 512, device=DEV)) for _ in range(n)]
    pc = tc.PackedConv(torch.randn(256, 256, 3, 3, device=DEV) * 0.03, torch.zeros(256, device=DEV))
    sl = torch.tensor([0.2], device=DEV)
    g, b = torch.ones(256, device=DEV), torch.zeros(256, device=DEV)
    us0 = bench(lambda i: tc.conv_tc(xin[i], pc, act=ops.ACT_PRELU, slope=sl), n)
    def sep(i):
        y = tc.conv_tc(xin[i], pc, act=ops.ACT_PRELU, slope=sl)
        return tc.batchnorm_c8(y, g, b, None, None, batch_stats=True, pool=True)
    def fused(i):
        y, part, mb = tc.conv_tc_bn_stats(xin[i], pc, act=ops.ACT_PRELU, slope=sl)
        return tc.batchnorm_c8(y, g, b, None, None, batch_stats=True, pool=True, partial=(part, mb))
    us1, us2 = bench(sep, n), bench(fused, n)
    if us0:
        print(f"conv 256->256 3x3 @512 + PReLU: {us0:.1f} us;  + BatchNorm (+pool) separate stats pass: {us1:.1f} us;  fused stats: {us2:.1f} us", flush=True)
elif what == "bn":
    n = 4
    c8s = [tc.to_c8(torch.randn(1, 256, 512, 512, device=DEV)) for _ in range(n)]
    g, b = torch.ones(256, device=DEV), torch.zeros(256, device=DEV)
    us = bench(lambda i: tc.batchnorm_c8(c8s[i], g, b, None, None, batch_stats=True, pool=True), n)
    if us:
        print(f"batchnorm_c8 256ch + pool: {us:.1f} us", flush=True)
torch.cuda.synchronize()
print("ok")

"""Times the two depth-stencil convolutions of the conditioning net (banded weights) at each level's depth count."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cwfa_b200 import tc, ops
DEV = "cuda:0"
for D in (48, 24, 12, 6):
    Cm = 32
    w1 = torch.randn(Cm, 3, 3, 3, device=DEV) * 0.2
    w2 = torch.randn(Cm, 3, 3, 3, device=DEV) * 0.05
    W1 = torch.zeros(D, Cm, D, 3, 3, device=DEV)
    W2 = torch.zeros(D, D, Cm, 3, 3, device=DEV)
    for kd in range(3):
        for d in range(D):
            dp = d + kd - 1
            if 0 <= dp < D:
                W1[d, :, dp] = w1[:, :, :, kd]
                W2[d, dp] = w2[:, :, :, kd]
    s1 = tc.PackedConv(W1.reshape(D * Cm, D, 3, 3), torch.zeros(D * Cm, device=DEV))
    s2 = tc.PackedConv(W2.reshape(D, D * Cm, 3, 3), torch.zeros(D, device=DEV))
    Wg = tc.col2im3x3_weights(W2.reshape(D, D * Cm, 3, 3))
    gp = tc.pad16(Wg.shape[0])
    variants = {}
    for bn in sorted({144 if gp % 144 == 0 else gp, gp if gp <= 256 else 144, 72 * 2 if gp % 144 == 0 else gp}):
        for mb in (1, 2):
            if bn * mb <= 512:
                variants[(bn, mb)] = tc.PackedConv(Wg, None, bn=bn)
    bias2 = torch.zeros(D, device=DEV)
    slope = torch.full((1,), 0.25, device=DEV)
    xs = [tc.to_c8(torch.randn(1, D, 512, 512, device=DEV)) for _ in range(4)]
    hs = [tc.conv_tc(x, s1, act=ops.ACT_PRELU, slope=slope) for x in xs]
    ys = [tc.conv_tc(h, s2) for h in hs]
    torch.cuda.synchronize()
    cases = [("s1", lambda i: tc.conv_tc(xs[i], s1, act=ops.ACT_PRELU, slope=slope)), ("s2 tap-by-tap", lambda i: tc.conv_tc(hs[i], s2))]
    if os.environ.get("S1_SWEEP"):
        for bn in (256, 192, 128, 64):
            if (D * Cm) % bn:
                continue
            pcv = tc.PackedConv(W1.reshape(D * Cm, D, 3, 3), torch.zeros(D * Cm, device=DEV), bn=bn)
            for mb in (1, 2):
                cases.append((f"s1 bn{bn} mb{mb}", lambda i, pcv=pcv, mb=mb: tc.conv_tc(xs[i], pcv, act=ops.ACT_PRELU, slope=slope, mb=mb)))
    for (bn, mb), pcg in variants.items():
        cases.append((f"s2 1x1 bn{bn} mb{mb}", lambda i, pcg=pcg, mb=mb: tc.conv_tc(hs[i], pcg, mb=mb)))
    g0 = tc.conv_tc(hs[0], next(iter(variants.values())))
    cases.append(("s2 col2im", lambda i: tc.col2im3x3_c8(g0, bias2, D)))
    for name, fn in cases:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep = [fn(i) for i in range(4)]
        g.replay(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 4)
        pc = s1 if name == "s1" else s2
        print(f"D={D} {name} BN={pc.BN} {best*1e3:.1f} us", flush=True)
        del keep, g

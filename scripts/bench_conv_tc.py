"""Micro-benchmark of the tcgen05 conv kernel on the shapes of the full 512x512x96 config."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cwfa_b200 import tc, ops

DEV = "cuda:0"
SHAPES = [  # cin, cout, k, H, W, mb, bn
    (64, 64, 3, 512, 512, 2, 64), (64, 64, 3, 512, 512, 1, 64), (64, 64, 1, 512, 512, 2, 64),
    (64, 96, 3, 512, 512, 2, 96), (48, 64, 1, 512, 512, 2, 64), (29, 48, 3, 512, 512, 2, 48),
    (16, 256, 3, 512, 512, 2, 256), (256, 256, 3, 512, 512, 2, 256), (256, 256, 3, 512, 512, 1, 256),
    (256, 512, 3, 256, 256, 2, 256), (512, 512, 3, 256, 256, 2, 256),
    (512, 1024, 3, 128, 128, 2, 256), (1024, 1024, 3, 128, 128, 2, 256), (1024, 1024, 3, 128, 128, 1, 256),
    (64, 64, 7, 512, 512, 2, 64),
    (256, 256, 3, 512, 512, 2, 128), (48, 1536, 3, 512, 512, 2, 256), (48, 1536, 3, 512, 512, 2, 128), (1536, 48, 3, 512, 512, 2, 48),
    (1536, 48, 3, 512, 512, 1, 48), (16, 256, 3, 512, 512, 2, 128), (512, 512, 3, 256, 256, 2, 128), (64, 96, 3, 512, 512, 2, 96), (64, 96, 3, 512, 512, 1, 96), (64, 48, 3, 512, 512, 2, 48),
    (512, 1024, 1, 256, 256, 2, 256), (1024, 2048, 1, 128, 128, 2, 256), (512, 1024, 1, 256, 256, 1, 128),
    (6, 6, 3, 512, 512, 2, 16), (6, 6, 7, 512, 512, 2, 16), (6, 6, 3, 512, 512, 1, 16),
]

def run(cin, cout, k, H, W, mb, bn, act=ops.ACT_ELU, reps=3):
    per = H * W * (tc.pad16(cin) + tc.pad16(cout)) * 2
    nbuf = max(2, min(64, (300 << 20) // per + 1))        # rotate through > L2 (126 MB) worth of tensors
    xs = [tc.to_c8(torch.randn(1, cin, H, W, device=DEV)) for _ in range(nbuf)]
    w = torch.randn(cout, cin, k, k, device=DEV) * (1.0 / (cin * k * k)) ** 0.5
    pc = tc.PackedConv(w, torch.zeros(cout, device=DEV), bn=bn)
    slope = torch.full((1,), 0.25, device=DEV) if act == ops.ACT_PRELU else None
    outs = [tc.conv_tc(x, pc, act=act, slope=slope, mb=mb) for x in xs]      # warm-up + keeps outputs alive (no allocator reuse)
    torch.cuda.synchronize()
    best = 1e9
    g = torch.cuda.CUDAGraph()          # graph replay: the host cost of a call (~50 us) must not bound the small shapes
    with torch.cuda.graph(g):
        for x in xs:
            tc.conv_tc(x, pc, act=act, slope=slope, mb=mb)
    g.replay()
    torch.cuda.synchronize()
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / nbuf)
    fl = 2.0 * H * W * tc.pad16(cin) * tc.pad16(cout) * k * k
    print(json.dumps(dict(shape=f"{cin}->{cout} k{k} {H}x{W} mb{mb} bn{bn}", us=round(best * 1e3, 1),
                          tflops=round(fl / best / 1e9, 1), gbs=round(per / best / 1e6, 1))), flush=True)


if __name__ == "__main__":
    sel = [int(a) for a in sys.argv[1:]] or range(len(SHAPES))
    for i in sel:
        run(*SHAPES[i], reps=4)

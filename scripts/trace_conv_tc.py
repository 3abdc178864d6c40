"""Per-CTA phase timeline of conv_tc (globaltimer stamps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cwfa_b200 import tc, ops, _lib
from bench_conv_tc import SHAPES
DEV = "cuda:0"
for i in [int(a) for a in sys.argv[1:]]:
    cin, cout, k, H, W, mb, bn = SHAPES[i]
    x = tc.to_c8(torch.randn(1, cin, H, W, device=DEV))
    pc = tc.PackedConv(torch.randn(cout, cin, k, k, device=DEV) * 0.05, torch.zeros(cout, device=DEV), bn=bn)
    for _ in range(2):
        tc.conv_tc(x, pc, act=ops.ACT_ELU, mb=mb)
    nct = min(296, ((W + 8 * mb - 1) // (8 * mb)) * ((H + 15) // 16) * (tc.pad16(cout) // bn))   # persistent grid: stamps = first item of each CTA
    dbg = torch.zeros(nct * 8, dtype=torch.int64, device=DEV)
    _lib.call("cwfa_tc_set_debug_buffer", dbg.data_ptr())
    tc.conv_tc(x, pc, act=ops.ACT_ELU, mb=mb)
    torch.cuda.synchronize()
    _lib.call("cwfa_tc_set_debug_buffer", None)
    d = dbg.view(nct, 8).cpu().double()
    t0 = d[:, 0].min()
    names = ["setup", "wait A", "wait B0", "issue MMAs", "MMA drain->acc_full", "epilogue", "final sync"]
    print(f"== {SHAPES[i]}  CTAs={nct}  span={(d[:,7].max()-t0)/1e3:.1f} us  mean CTA life={(d[:,7]-d[:,0]).mean()/1e3:.2f} us")
    for j, nm in enumerate(names):
        print(f"   {nm:22s} mean {(d[:, j+1]-d[:, j]).mean()/1e3:8.2f} us   max {(d[:, j+1]-d[:, j]).max()/1e3:8.2f}")
    starts = (d[:, 0] - t0).sort().values / 1e3
    print("   CTA start times (us) deciles:", [round(float(starts[int(q * (nct - 1) / 10)]), 1) for q in range(11)])

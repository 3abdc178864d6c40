import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cwfa_b200 import tc, _lib
DEV = "cuda:0"
for ch in (6, 48):
    b = tc.to_c8(torch.randn(1, 64, 512, 512, device=DEV))
    pc = tc.PackedConv(torch.randn(2 * ch, 64, 3, 3, device=DEV) * 0.04, torch.zeros(2 * ch, device=DEV), bn=tc.pad16(2 * ch))
    x = torch.randn(1, ch, 512, 512, device=DEV)
    perm = torch.randperm(ch, device=DEV).to(torch.int32)
    ld = torch.zeros(1, device=DEV)
    for _ in range(3):
        tc.conv_tc_coupling(b, pc, x, ch=ch, inverse=True, perm=perm, perm_axis=1, logdet=ld)
    nct = 1024
    dbg = torch.zeros(nct * 8, dtype=torch.int64, device=DEV)
    _lib.call("cwfa_tc_set_debug_buffer", dbg.data_ptr())
    tc.conv_tc_coupling(b, pc, x, ch=ch, inverse=True, perm=perm, perm_axis=1, logdet=ld)
    torch.cuda.synchronize()
    _lib.call("cwfa_tc_set_debug_buffer", None)
    d = dbg.view(nct, 8).cpu().double()
    t0 = d[:, 0].min()
    names = ["setup", "wait A", "wait B0", "issue MMAs", "MMA drain->acc_full", "epilogue", "final sync"]
    print(f"== ch={ch} span={(d[:,7].max()-t0)/1e3:.1f} us  mean CTA life={(d[:,7]-d[:,0]).mean()/1e3:.2f} us")
    for j, nm in enumerate(names):
        print(f"   {nm:22s} mean {(d[:, j+1]-d[:, j]).mean()/1e3:8.2f} us   max {(d[:, j+1]-d[:, j]).max()/1e3:8.2f}")
    starts = (d[:, 0] - t0).sort().values / 1e3
    print("   CTA start deciles:", [round(float(starts[int(q * (nct - 1) / 10)]), 1) for q in range(11)])

"""Per-shape time of every C-ABI call in one eager reconstruct step (CUDA events around each call)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cwfa_b200
from cwfa_b200 import _lib, tc, ops
from cwfa_b200.engine import CWFAEngine
from bench import synthetic_inputs
side = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfg = dict(side=side, depths=96, steps=5)
dev = torch.device("cuda:0")
model = cwfa_b200.CWFAModel(n_depths=96, volume_side_size=side, INN_max_down_steps=5, seed=0).to(dev)
eng = CWFAEngine(model, "bf16")
views, mvs = synthetic_inputs(cfg, dev, 100)
views, mvs = views.to(dev), [m.to(dev) for m in mvs]
for _ in range(2):
    eng.reconstruct(views, mvs)
torch.cuda.synchronize()
evs = []
orig = _lib.call
def timed(name, *a):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); orig(name, *a); e.record()
    key = name
    if name == "cwfa_conv_tc":
        N, H, W, Cin_p, Cout, Cout_p, KH, KW, BN, MB, act, res_mode, out_mode = a[6:19]
        key = f"conv_tc {Cin_p:4d}->{Cout_p:4d} k{KH} {H}x{W} BN{BN} MB{MB} out{out_mode} res{res_mode}"
        fl = 2.0 * H * W * Cin_p * Cout_p * KH * KW * (4 if out_mode == 2 else 1)
    elif name == "cwfa_coupling_tc":
        N, H, W, Cout, Cout_p = a[3:8]
        key = f"coupling_tc 64->{Cout_p} ch{a[14]} axis{a[13]} x{'1' if a[8] else '0'}"
        fl = 2.0 * H * W * 64 * Cout_p * 9
    elif name == "cwfa_conv_tc_coupling":
        N, H, W, Cin_p, Cout, Cout_p, KH, KW, MB = a[3:12]
        key = f"conv_tc_coupling {Cin_p}->{Cout_p} k{KH} ch{a[18]} axis{a[17]} x{'1' if a[12] else '0'}"
        fl = 2.0 * H * W * Cin_p * Cout_p * KH * KW
    else:
        fl = 0
    evs.append((key, s, e, fl))
for mod in (_lib, tc, ops):
    if hasattr(mod, "_lib"): mod._lib.call = timed
_lib.call = timed
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.reconstruct(views, mvs); e1.record()
torch.cuda.synchronize()
agg = collections.OrderedDict()
for k, s, e, fl in evs:
    a = agg.setdefault(k, [0, 0.0, 0.0]); a[0] += 1; a[1] += s.elapsed_time(e); a[2] += fl
tot = sum(a[1] for a in agg.values())
print(f"step wall (eager, with events) {e0.elapsed_time(e1):.2f} ms; sum of calls {tot:.2f} ms")
for k, (c, t, fl) in sorted(agg.items(), key=lambda x: -x[1][1]):
    extra = f"  {fl / t / 1e9:7.1f} TFLOP/s(padded)" if fl else ""
    print(f"{t*1e3:9.1f} us {100*t/tot:5.1f}% x{c:3d}  {k}{extra}")

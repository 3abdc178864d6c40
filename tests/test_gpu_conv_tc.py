"""GPU: the tcgen05 implicit-GEMM convolution vs torch CPU convolution on operands rounded to the same
half-precision grid (so the only difference is fp32 accumulation order)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import max_abs, rel_l2
from oracle.weights import seeded_randn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _round(t, kind):
    return t.to(torch.bfloat16 if kind == "bf16" else torch.float16).float()


CASES = [
    # cin, cout, k, H, W, mb, bn
    (64, 64, 3, 32, 32, 1, 64),
    (64, 64, 3, 32, 32, 2, 64),
    (64, 64, 1, 16, 24, 2, 64),
    (16, 16, 3, 16, 16, 1, 16),
    (6, 64, 1, 20, 13, 1, 64),
    (48, 96, 3, 40, 24, 2, 96),
    (29, 48, 3, 33, 17, 2, 48),
    (64, 12, 3, 16, 16, 2, 16),
    (128, 64, 3, 32, 16, 2, 64),
    (256, 256, 3, 32, 32, 2, 256),
    (64, 64, 7, 32, 32, 2, 64),
    (256, 512, 3, 16, 16, 1, 256),
]


@pytest.mark.parametrize("cin,cout,k,H,W,mb,bn", CASES)
@pytest.mark.parametrize("kind", ["bf16", "fp16"])
def test_conv_tc_vs_cpu(cin, cout, k, H, W, mb, bn, kind):
    from cwfa_b200 import ops, tc
    x = _round(seeded_randn((2, cin, H, W), 1), kind)
    w = _round(seeded_randn((cout, cin, k, k), 2, (1.0 / (cin * k * k)) ** 0.5), kind)
    b = seeded_randn((cout,), 3)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=k // 2).float()
    xc = tc.to_c8(x.to(DEV), kind)
    assert torch.equal(tc.from_c8(xc).cpu(), x), "C8 round trip must be exact on representable values"
    pc = tc.PackedConv(w.to(DEV), b.to(DEV), kind, bn=bn)
    y = tc.conv_tc(xc, pc, out_nchw=True, mb=mb)
    torch.cuda.synchronize()
    err = rel_l2(y, ref)
    print(f"conv_tc {cin}->{cout} k{k} {H}x{W} mb{mb} bn{bn} {kind}: rel_l2={err:.2e} max_abs={max_abs(y, ref):.2e}")
    assert err < 2e-5
    # C8 output with fused ELU + residual (pre-activation add), rounded to half on store
    r = _round(seeded_randn((2, cout, H, W), 4), kind)
    y2 = tc.from_c8(tc.conv_tc(xc, pc, act=ops.ACT_ELU, res=tc.to_c8(r.to(DEV), kind), res_mode=1, mb=mb))
    ref2 = F.elu(ref + r)
    tol = 6e-3 if kind == "bf16" else 8e-4
    assert rel_l2(y2, ref2) < tol


@pytest.mark.parametrize("kind", ["bf16", "fp16"])
@pytest.mark.parametrize("N,H,W", [(1, 16, 16), (2, 48, 32), (1, 40, 24), (1, 128, 128), (3, 64, 80)])
def test_fused_resblock_vs_unfused_and_cpu(kind, N, H, W):
    """Fused persistent trunk block == the two-kernel sequence (same roundings) and ~= fp64 CPU."""
    from cwfa_b200 import ops, tc
    x = _round(seeded_randn((N, 64, H, W), 11), kind)
    w3 = _round(seeded_randn((64, 64, 3, 3), 12, (1.0 / 576) ** 0.5), kind)
    w1 = _round(seeded_randn((64, 64, 1, 1), 13, (1.0 / 64) ** 0.5), kind)
    b3, b1 = seeded_randn((64,), 14, 0.1), seeded_randn((64,), 15, 0.1)
    xc = tc.to_c8(x.to(DEV), kind)
    p3, p1 = tc.PackedConv(w3.to(DEV), b3.to(DEV), kind, bn=64), tc.PackedConv(w1.to(DEV), b1.to(DEV), kind, bn=64)
    y = tc.from_c8(tc.resblock_tc(xc, p3, p1))
    t = tc.conv_tc(xc, p3, act=ops.ACT_ELU)
    y_ref2 = tc.from_c8(tc.conv_tc(t, p1, act=ops.ACT_ELU, res=xc, res_mode=1))
    torch.cuda.synchronize()
    tmid = _round(F.elu(F.conv2d(x.double(), w3.double(), b3.double(), padding=1)).float(), kind)
    ref = F.elu(F.conv2d(tmid.double(), w1.double(), b1.double()) + x.double()).float()
    tol = 6e-3 if kind == "bf16" else 8e-4
    print(f"resblock {kind} {N}x{H}x{W}: vs unfused {rel_l2(y, y_ref2):.2e}  vs cpu {rel_l2(y, ref):.2e}")
    assert rel_l2(y, y_ref2) < tol
    assert rel_l2(y, ref) < tol

"""GPU: the fused true-3-D depth-stencil kernel (csrc/stencil_tc.cu) against torch's Conv3d -> PReLU -> Conv3d -- the reference's
ResidualBlock.conv3d (networks.py:221-225 applied at :239; Dropout3d is the identity in eval mode) -- and against the banded
two-convolution form it replaces in the engine."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _weights(seed, slope):
    g = torch.Generator().manual_seed(seed)
    w1 = (torch.randn(32, 1, 3, 3, 3, generator=g) * 0.3).to(DEV)
    b1 = (torch.randn(32, generator=g) * 0.2).to(DEV)
    w2 = (torch.randn(1, 32, 3, 3, 3, generator=g) * 0.1).to(DEV)
    b2 = (torch.randn(1, generator=g) * 0.2).to(DEV)
    return w1, b1, w2, b2, torch.tensor([slope], device=DEV)


def _reference(x, w1, b1, w2, b2, slope, dt=None):
    """x (N, D, H, W) fp32.  ``dt``: round the operands of both convolutions to that half type (what the tensor cores see)."""
    r = (lambda t: t.to(dt).float()) if dt is not None else (lambda t: t)
    v = r(x).permute(0, 2, 3, 1).unsqueeze(1)                                   # (N, 1, H, W, D), networks.py:239
    h = F.prelu(F.conv3d(v, r(w1), b1, padding=1), slope)
    o = F.conv3d(r(h), r(w2), b2, padding=1)
    return o[:, 0].permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("kind", ["bf16", "fp16"])
@pytest.mark.parametrize("N,D,H,W,slope", [(1, 8, 64, 64, 0.25), (2, 48, 24, 40, 0.25), (1, 24, 37, 150, 1.5), (1, 12, 16, 300, -0.3),
                                           (1, 6, 33, 512, 0.1), (1, 5, 9, 7, 0.25), (1, 1, 12, 20, 0.25), (1, 64, 5, 30, 0.25)])
def test_stencil3d_tc_matches_conv3d(kind, N, D, H, W, slope):
    from cwfa_b200 import tc
    torch.backends.cudnn.allow_tf32 = False
    dt = torch.bfloat16 if kind == "bf16" else torch.float16
    w1, b1, w2, b2, sl = _weights(3 + D, slope)
    x = torch.randn(N, D, H, W, generator=torch.Generator().manual_seed(D + H)).to(DEV)
    sw = tc.StencilWeights(w1, b1, w2, b2, kind)
    y8 = tc.stencil3d_tc(tc.to_c8(x, kind), sw, sl, D)
    assert y8.Cp == tc.pad16(D)
    got = tc.from_c8(y8)
    assert torch.isfinite(y8.data.float()).all()
    if y8.Cp > D:                                                              # channel padding is written as zeros
        full = y8.data.float().permute(0, 1, 4, 2, 3).reshape(N, y8.Cp, H, W)
        assert float(full[:, D:].abs().max()) == 0.0
    emu = _reference(x, w1, b1, w2, b2, sl, dt)
    exact = _reference(x, w1, b1, w2, b2, sl)
    e_emu, e_exact = rel_l2(got, emu.to(dt).float()), rel_l2(got, exact)
    print(f"stencil3d_tc {kind} D={D} {H}x{W}: rel-L2 vs half-operand emulation {e_emu:.2e}, vs fp32 {e_exact:.2e}")
    assert e_emu < (4e-3 if kind == "bf16" else 1.2e-3)
    assert e_exact < (1.5e-2 if kind == "bf16" else 2e-3)


@pytest.mark.parametrize("rows_max", [128, 384, 768])
def test_stencil3d_tc_strip_geometry_is_invisible(rows_max):
    """Different strip widths (GEMM rows per pixel-row) give the same result bit for bit."""
    from cwfa_b200 import tc
    w1, b1, w2, b2, sl = _weights(11, 0.25)
    x = torch.randn(1, 12, 50, 90, generator=torch.Generator().manual_seed(5)).to(DEV)
    sw = tc.StencilWeights(w1, b1, w2, b2, "bf16")
    x8 = tc.to_c8(x, "bf16")
    a = tc.stencil3d_tc(x8, sw, sl, 12).data
    b = tc.stencil3d_tc(x8, sw, sl, 12, rows_max=rows_max).data
    assert torch.equal(a, b)


def test_stencil3d_tc_agrees_with_banded_form_full_size():
    """512 x 512 x 48 (flow level 0): the fused kernel against the banded s1 -> s2 -> col2im form (both bf16 operand paths)."""
    from cwfa_b200 import tc, ops
    D = 48
    w1, b1, w2, b2, sl = _weights(7, 0.25)
    x = torch.randn(1, D, 512, 512, generator=torch.Generator().manual_seed(1)).to(DEV)
    x8 = tc.to_c8(x, "bf16")
    got = tc.from_c8(tc.stencil3d_tc(x8, tc.StencilWeights(w1, b1, w2, b2, "bf16"), sl, D))
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
    s2g = tc.PackedConv(Wg, None, "bf16", bn=144)
    hid = tc.conv_tc(x8, s1, act=ops.ACT_PRELU, slope=sl)
    ref = tc.from_c8(tc.col2im3x3_c8(tc.conv_tc(hid, s2g, mb=1), b2.repeat(D), D))
    exact = _reference(x, w1, b1, w2, b2, sl)
    e_f, e_b = rel_l2(got, exact), rel_l2(ref, exact)
    print(f"level-0 stencil vs fp32 Conv3d: fused {e_f:.2e}, banded {e_b:.2e}; fused vs banded {rel_l2(got, ref):.2e}")
    assert e_f < 1.5e-2 and e_f < 1.5 * e_b + 1e-3

"""GPU: every module of the boundary vs (a) golden outputs of the unmodified reference and (b) the
CPU oracle on seeded inputs; round trips and log-det sign convention (FrEIA/modules/base.py:26-31)."""
import pytest
import torch

from conftest import max_abs, rel_l2
from oracle import cwfa_oracle as O
from oracle.weights import deterministic_fill, seeded_randn
from test_oracle_golden import _block_module, block_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 5e-5       # fp32 kernels vs fp32 reference (different summation order inside the convolutions)


def test_library_loaded_and_device_is_b200():
    from cwfa_b200 import _lib
    _lib.call("cwfa_device_check")


def test_haar1d(golden_modules):
    import cwfa_b200.modules as Fm
    x = seeded_randn((2, 12, 10, 14), 20)
    m = Fm.HaarTransform1D([(12, 10, 14)], order_by_wavelet=True)
    (y,), j = m((x.to(DEV),), rev=False)
    assert max_abs(y, golden_modules["haar1d/fwd"]) < 1e-6 and j == 0.0
    (xr,), jr = m((y,), rev=True)
    assert max_abs(xr, x) < 1e-6 and jr == 0.0
    (r,), _ = m((x.to(DEV),), rev=True)
    assert max_abs(r, golden_modules["haar1d/rev_of_x"]) < 1e-6


@pytest.mark.parametrize("shape", [(1, 2, 1, 1), (3, 6, 5, 7), (1, 96, 64, 64), (2, 4, 3, 4)])
def test_haar1d_shapes_vs_oracle(shape):
    from cwfa_b200 import ops
    x = seeded_randn(shape, 5)
    y = ops.haar1d_forward(x.to(DEV))
    assert max_abs(y, O.haar1d(x)[0]) < 1e-6
    lo, hi = ops.haar1d_split(x.to(DEV))
    assert torch.equal(torch.cat([lo, hi], 1), y)
    assert max_abs(ops.haar1d_merge(lo, hi), x) < 1e-6
    assert max_abs(ops.haar1d_inverse(y), x) < 1e-6


def test_haar2d(golden_modules):
    import cwfa_b200.modules as Fm
    x = seeded_randn((2, 3, 8, 12), 21)
    for obw in (False, True):
        for reb in (1.0, 0.5):
            m = Fm.HaarDownsampling([(3, 8, 12)], order_by_wavelet=obw, rebalance=reb)
            xin = x.to(DEV)
            (y,), j = m((xin,), rev=False)
            assert max_abs(y, golden_modules[f"haar2d/down/{int(obw)}/{reb}"]) < 1e-6
            assert abs(j - float(golden_modules[f"haar2d/down_jac/{int(obw)}/{reb}"])) < 1e-3
            yin = y.clone()
            (xr,), jr = m((yin,), rev=True)
            assert torch.equal(yin, y), "input must not be mutated (reference bug reshapes.py:297 not replicated)"
            assert max_abs(xr, golden_modules[f"haar2d/up/{int(obw)}/{reb}"]) < 1e-6
            assert abs(jr - float(golden_modules[f"haar2d/up_jac/{int(obw)}/{reb}"])) < 1e-3
            up = Fm.HaarUpsampling([(12, 4, 6)], order_by_wavelet=obw, rebalance=reb)
            (xu,), _ = up((y,), rev=False)
            assert max_abs(xu, xr) == 0.0


def test_permutations(golden_modules):
    import numpy as np
    import cwfa_b200.modules as Fm
    x = seeded_randn((2, 6, 8, 8), 22).to(DEV)
    m = Fm.PermuteRandom([(6, 8, 8)], seed=3)
    assert torch.equal(m.perm.data, golden_modules["perm_chan/perm"]), "numpy-seeded permutation must match the reference"
    assert torch.equal(m((x,))[0][0].cpu(), golden_modules["perm_chan/fwd"])
    assert torch.equal(m((x,), rev=True)[0][0].cpu(), golden_modules["perm_chan/rev"])
    for t in range(4):
        ax = int(golden_modules[f"perm_dim/{t}/axis"])
        m = Fm.PermuteDim([(6, 8, 8)], seed=5 + t, axis=ax)
        assert torch.equal(m.perm.data, golden_modules[f"perm_dim/{t}/perm"])
        assert torch.equal(m((x,))[0][0].cpu(), golden_modules[f"perm_dim/{t}/fwd"])
        assert torch.equal(m((x,), rev=True)[0][0].cpu(), golden_modules[f"perm_dim/{t}/rev"])
        assert torch.equal(m((m((x,))[0][0],), rev=True)[0][0], x)


@pytest.mark.parametrize("name", ["cat", "cat_first", "GLOW", "GIN", "RNVP"])
def test_blocks_vs_reference_golden(golden_modules, name):
    m = _block_module(name).to(DEV)
    x, c_lf, c_mv = (t.to(DEV) for t in block_inputs())
    conds = [c_mv, c_lf] if name == "cat_first" else [c_lf]
    (y,), j = m((x,), c=conds, rev=False)
    (xr,), jr = m((x,), c=conds, rev=True)
    assert rel_l2(y, golden_modules[f"block/{name}/fwd"]) < TOL
    assert rel_l2(xr, golden_modules[f"block/{name}/rev"]) < TOL
    if name != "GIN":
        assert max_abs(j, golden_modules[f"block/{name}/fwd_jac"]) < 5e-3
        assert max_abs(jr, golden_modules[f"block/{name}/rev_jac"]) < 5e-3
    # invertibility and sign convention
    (x2,), j2 = m((y,), c=conds, rev=True)
    assert rel_l2(x2, x) < 1e-5
    if name != "GIN":
        assert max_abs(j2, -j) < 5e-3


def test_affine_z_none_equals_zero_input():
    from cwfa_b200 import ops
    a = seeded_randn((2, 8, 6, 10), 7).to(DEV)
    x0 = torch.zeros(2, 4, 6, 10, device=DEV)
    y0, j0 = ops.affine(x0, a[:, :4], a[:, 4:], inverse=True)
    y1, j1 = ops.affine(None, a[:, :4], a[:, 4:], inverse=True)
    assert torch.equal(y0, y1) and torch.equal(j0, j1)
    # deterministic two-stage reduction: bit-identical on repeat
    y2, j2 = ops.affine(None, a[:, :4], a[:, 4:], inverse=True)
    assert torch.equal(j1, j2)
    yo, jo = O.affine(x0.cpu(), a.cpu(), rev=True)
    assert rel_l2(y1, yo) < 1e-6 and max_abs(j1, jo) < 1e-3


@pytest.mark.parametrize("cin,cout,k,hw", [(3, 5, 3, (9, 11)), (29, 6, 1, (16, 16)), (6, 6, 7, (20, 33)),
                                           (64, 64, 3, (32, 32)), (16, 40, 3, (17, 5))])
def test_conv2d_f32_vs_torch_cpu(cin, cout, k, hw):
    import torch.nn.functional as F
    from cwfa_b200 import ops
    x = seeded_randn((2, cin) + hw, 1)
    w = seeded_randn((cout, cin, k, k), 2, 0.2)
    b = seeded_randn((cout,), 3)
    r = seeded_randn((2, cout) + hw, 4)
    ref = F.elu(F.conv2d(x, w, b, padding=k // 2) + r)
    y = ops.conv2d(x.to(DEV), w.to(DEV), b.to(DEV), act=ops.ACT_ELU, res=r.to(DEV), res_mode=1)
    assert rel_l2(y, ref) < 1e-5
    ref2 = F.gelu(F.conv2d(x, w, None, padding=k // 2)) + r
    y2 = ops.conv2d(x.to(DEV), w.to(DEV), None, act=ops.ACT_GELU, res=r.to(DEV), res_mode=2)
    assert rel_l2(y2, ref2) < 1e-5


def test_cond_network_vs_golden(golden_tiny):
    from helpers import build_tiny_model, tiny_inputs
    m = build_tiny_model(golden_tiny, DEV)
    views, _ = tiny_inputs(golden_tiny)
    for n in range(m.n_levels):
        out = m.cond_nets[n](views.to(DEV))[-1]
        assert rel_l2(out, golden_tiny[f"cond{n}"]) < TOL


@pytest.mark.parametrize("ch,H,W", [(6, 13, 9), (48, 16, 24), (3, 8, 8)])
def test_depth_stencil_vs_torch_cpu(ch, H, W):
    import torch.nn.functional as F
    from cwfa_b200 import ops
    x = seeded_randn((2, ch, H, W), 1)
    w1, b1 = seeded_randn((32, 1, 3, 3, 3), 2, 0.3), seeded_randn((32,), 3, 0.1)
    w2, b2 = seeded_randn((1, 32, 3, 3, 3), 4, 0.1), seeded_randn((1,), 5, 0.1)
    a = torch.tensor([0.2])
    v = x.permute(0, 2, 3, 1).unsqueeze(1)
    ref = F.conv3d(F.prelu(F.conv3d(v, w1, b1, padding=1), a), w2, b2, padding=1)[:, 0].permute(0, 3, 1, 2)
    y = ops.depth_stencil3d(x.to(DEV), w1.to(DEV), b1.to(DEV), a.to(DEV), w2.to(DEV), b2.to(DEV))
    assert rel_l2(y, ref) < 1e-5


def test_batchnorm_train_mode_updates_running_statistics_like_torch():
    """unet.py:100-107 uses nn.BatchNorm2d and the reference leaves its LRNN in .train() mode (CWFA.py:531-532): every forward
    normalises with batch statistics AND moves running_mean / running_var (unbiased variance, momentum 0.1) and
    num_batches_tracked -- the values its checkpoints carry.  Same side effect here, checked against torch's own module."""
    import torch.nn as nn
    from cwfa_b200 import ops
    C = 8
    ref = nn.BatchNorm2d(C)
    with torch.no_grad():
        ref.weight.copy_(seeded_randn((C,), 1) * 0.3 + 1.0)
        ref.bias.copy_(seeded_randn((C,), 2) * 0.2)
        ref.running_mean.copy_(seeded_randn((C,), 3) * 0.1)
        ref.running_var.copy_(seeded_randn((C,), 4).abs() + 0.5)
    ours = nn.BatchNorm2d(C)
    ours.load_state_dict(ref.state_dict())
    ours = ours.to(DEV)
    for step in range(3):
        x = seeded_randn((2, C, 9, 11), 10 + step) * (1.0 + step) + 0.3 * step
        ref.train()
        y_ref = ref(x)
        y = ops.batchnorm(x.to(DEV), ours.weight, ours.bias, ours.running_mean, ours.running_var, batch_stats=True, eps=ours.eps,
                          momentum=ours.momentum, update_running=True, num_batches_tracked=ours.num_batches_tracked)
        assert rel_l2(y, y_ref.detach()) < 1e-5
        assert max_abs(ours.running_mean, ref.running_mean) < 1e-6 and rel_l2(ours.running_var, ref.running_var) < 1e-6
        assert int(ours.num_batches_tracked) == int(ref.num_batches_tracked) == step + 1
    ref.eval()
    x = seeded_randn((2, C, 9, 11), 20)
    y = ops.batchnorm(x.to(DEV), ours.weight, ours.bias, ours.running_mean, ours.running_var, batch_stats=False, eps=ours.eps)
    assert rel_l2(y, ref(x).detach()) < 1e-5
    # the differentiable path (training step) applies the same update
    xg = (seeded_randn((2, C, 9, 11), 30)).to(DEV).requires_grad_(True)
    ref.train()
    ref(seeded_randn((2, C, 9, 11), 30))
    ops.batchnorm(xg, ours.weight, ours.bias, ours.running_mean, ours.running_var, batch_stats=True, eps=ours.eps, momentum=ours.momentum,
                  update_running=True, num_batches_tracked=ours.num_batches_tracked).sum().backward()
    assert max_abs(ours.running_mean, ref.running_mean) < 1e-6 and int(ours.num_batches_tracked) == 4


@pytest.mark.parametrize("B,C,H,W", [(1, 6, 40, 512), (2, 3, 17, 36), (1, 5, 9, 10), (1, 2, 33, 7)])
def test_permute_gathers_bit_exact_on_every_axis(B, C, H, W):
    """K4 gather kernels (16-byte paths when W % 4 == 0, scalar otherwise) against index_select: PermuteRandom is axis 1
    (fixed_transforms.py:37-41), PermuteDim axes 2 / 3 (INN_utils.py:73-81)."""
    from cwfa_b200 import ops
    g = torch.Generator().manual_seed(B * 100 + W)
    x = torch.randn(B, C, H, W, generator=g).to(DEV)
    for axis, n in ((1, C), (2, H), (3, W)):
        perm = torch.randperm(n, generator=g)
        y = ops.permute(x, perm, axis)
        assert torch.equal(y, x.index_select(axis, perm.to(DEV)))

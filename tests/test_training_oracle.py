"""CPU: pins the oracle's TRAINING-step restatement (oracle/cwfa_oracle.py:level_train_loss / level_train_grads) against the
gradients the unmodified reference modules produce under torch autograd (tests/golden/train.pt, made by
tests/golden/make_golden_train.py), and checks the host-side logic of cwfa_b200/training.py (flat parameter groups).
No kernels are launched."""
import os
import sys

import pytest
import torch

from conftest import GOLDEN
from helpers import build_tiny_model

sys.path.insert(0, GOLDEN)
from oracle import cwfa_oracle as O                      # noqa: E402
from oracle.weights import seeded_randn                  # noqa: E402


@pytest.fixture(scope="module")
def golden_train():
    return torch.load(os.path.join(GOLDEN, "train.pt"), weights_only=False)


def probe(shape, key):
    seed = (sum(ord(c) * (i + 1) for i, c in enumerate(key)) % 100000) + 7
    return seeded_randn(tuple(shape), seed)


def train_inputs(fx, n):
    cfg = fx["config"]
    D, S, B = cfg["D"], cfg["S"], cfg["B"]
    C = D // 2 ** n
    views = seeded_randn((B, 29, S, S), cfg["seeds"]["views"])
    gt = seeded_randn((B, C, S, S), cfg["seeds"]["gt"] + 10 * n)
    vol_in = seeded_randn((B, C // 2, S, S), cfg["seeds"]["vol_in"] + 10 * n)
    mean_vol = seeded_randn((1, C // 2, S, S), cfg["seeds"]["mean"] + n, 0.1).repeat(B, 1, 1, 1)
    return gt, views, mean_vol, vol_in


def check_against_golden(grads, gold, tol):
    """grads: key -> tensor; gold: key -> {norm, probe, full?} from the reference."""
    assert set(gold) <= set(grads)
    for k, e in gold.items():
        g = grads[k].detach().double().cpu()
        scale = max(e["norm"], 1e-12)
        assert abs(g.norm().item() - e["norm"]) <= tol * scale, (k, g.norm().item(), e["norm"])
        pr = (g * probe(g.shape, k).double()).sum().item()
        assert abs(pr - e["probe"]) <= tol * scale * g.numel() ** 0.5, (k, pr, e["probe"])
        if "full" in e:
            assert (g - e["full"].double()).norm().item() <= tol * scale, k


@pytest.mark.parametrize("n", [0, 1])
def test_oracle_train_grads_match_reference_autograd(golden_tiny, golden_train, n):
    model = build_tiny_model(golden_tiny).export_for_oracle()
    lv = model["levels"][n]
    gt, views, mean_vol, vol_in = train_inputs(golden_train, n)
    r = O.level_train_grads(lv["inn"], lv["cond"], lv["spec"], gt, views, mean_vol, vol_in, golden_train["config"]["cond_weight"])
    g = golden_train[f"level{n}"]
    for key in ("loss", "mse", "nll"):
        assert abs(float(r[key]) - float(g[key])) <= 2e-5 * abs(float(g[key])), key
    check_against_golden(r["inn"], g["inn"], 2e-4)
    check_against_golden(r["cond"], g["cond"], 2e-4)
    # parameters the reference leaves without a gradient (unused weights kept for checkpoints) are zero in the oracle too
    for k in g["no_grad_keys"]:
        t = r["inn"].get(k, r["cond"].get(k))
        assert t is not None and float(t.abs().max()) == 0.0, k


def test_oracle_lrnn_grads_match_reference_autograd(golden_tiny, golden_train):
    """LRNN step: F.mse_loss(gt, Encoder(views)) through the U-Net (BatchNorm batch statistics, max-pool, transposed conv)."""
    model = build_tiny_model(golden_tiny).export_for_oracle()
    cfg, g = golden_train["config"], golden_train["lrnn"]
    D, S, B, MAX = cfg["D"], cfg["S"], cfg["B"], cfg["MAX"]
    views = seeded_randn((B, 29, S, S), g["seeds"]["views"])
    gt = seeded_randn((B, D // 2 ** (MAX - 1), S, S), g["seeds"]["gt"])
    r = O.lrnn_train_grads(model["lrnn"], views, gt)
    assert abs(float(r["loss"]) - float(g["loss"])) <= 2e-5 * abs(float(g["loss"]))
    check_against_golden(r["grads"], g["grads"], 3e-4)
    assert all(k not in r["grads"] for k in g["no_grad_keys"])          # mean-volume branch: unused without a mean volume


def test_oracle_lrnn_grads_with_mean_volume_branch(golden_tiny, golden_train):
    """LRNN step incl. the mean-volume branch (ConvNeXt 7x7 + LayerNorm + GELU, attention gate): all 79 parameters get gradients."""
    model = build_tiny_model(golden_tiny).export_for_oracle()
    cfg, g = golden_train["config"], golden_train["lrnn_mv"]
    D, S, B, MAX = cfg["D"], cfg["S"], cfg["B"], cfg["MAX"]
    nd = D // 2 ** (MAX - 1)
    views, gt = seeded_randn((B, 29, S, S), g["seeds"]["views"]), seeded_randn((B, nd, S, S), g["seeds"]["gt"])
    mv = seeded_randn((B, nd, S, S), g["seeds"]["mean_vol"], 0.1)
    r = O.lrnn_train_grads(model["lrnn"], views, gt, mv)
    assert abs(float(r["loss"]) - float(g["loss"])) <= 2e-5 * abs(float(g["loss"]))
    assert set(r["grads"]) == set(g["grads"]) and not g["no_grad_keys"]
    check_against_golden(r["grads"], g["grads"], 3e-4)


def test_flat_group_rehomes_parameters():
    from cwfa_b200.training import FlatGroup
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 5, 3), torch.nn.PReLU(), torch.nn.Conv2d(5, 2, 1))
    before = [p.detach().clone() for p in net.parameters()]
    fg = FlatGroup(net.parameters())
    assert all(torch.equal(a, p.detach()) for a, p in zip(before, net.parameters()))
    assert all(o % 4 == 0 for o in fg.offsets)
    assert all(p.data_ptr() == fg.flat.data_ptr() + 4 * o for p, o in zip(fg.params, fg.offsets))
    x = torch.randn(2, 3, 8, 8)
    net(x).square().sum().backward()                      # plain torch autograd: accumulates INTO the flat gradient views
    assert fg.grads_alias_flat()
    ref = torch.cat([torch.nn.functional.pad(p.grad.reshape(-1), (0, (-p.numel()) % 4)) for p in fg.params])
    assert torch.equal(ref, fg.grad) and float(fg.grad.abs().sum()) > 0
    # a parameter claimed by a second group stays "loose" there (the conditioning nets share ONE PReLU, networks.py:209)
    fg2 = FlatGroup([net[1].weight, torch.nn.Parameter(torch.zeros(3))])
    assert len(fg2.loose) == 1 and len(fg2.params) == 1
    assert all(fg.used) and fg.mask() is None             # every parameter received a gradient (autograd hooks)
    fg.zero_grad()
    assert float(fg.grad.abs().sum()) == 0.0 and not any(fg.used)
    net[0](x).sum().backward()                            # only the first conv is used now
    assert fg.used == [True, True, False, False, False]
    m = fg.mask()
    assert int(m.sum()) == net[0].weight.numel() + net[0].bias.numel() and int(m[: net[0].weight.numel()].min()) == 1


def test_oracle_lion_matches_published_update():
    p, g, m = torch.tensor([1.0, -2.0, 0.5, 3.0]), torch.tensor([0.3, -0.1, 0.0, 2.0]), torch.tensor([-1.0, 0.2, 0.0, -0.1])
    p2, m2 = O.lion_step(p, g, m, lr=0.1, beta1=0.9, beta2=0.99, wd=0.01)
    u = torch.sign(0.9 * m + 0.1 * g)                     # [-1, +1, 0, +1]
    assert torch.equal(u, torch.tensor([-1.0, 1.0, 0.0, 1.0]))
    assert torch.allclose(p2, p * (1 - 0.1 * 0.01) - 0.1 * u) and torch.allclose(m2, 0.99 * m + 0.01 * g)


def test_banded_depth_stencil_weights_equal_the_conv3d():
    """The depth-banded 2-D form of the conditioning net's Conv3d pair (cwfa_b200/autograd.py:depth_stencil3d_banded, also the
    inference engine's formulation) reproduces conv3d -> prelu -> conv3d of the reference (networks.py:221-225,236-241) on CPU."""
    import torch.nn.functional as F
    from cwfa_b200 import autograd as ag
    D, Cm, H, W = 6, 5, 7, 9
    x = seeded_randn((2, D, H, W), 1)
    w1, b1 = seeded_randn((Cm, 1, 3, 3, 3), 2, 0.3), seeded_randn((Cm,), 3)
    w2, b2 = seeded_randn((1, Cm, 3, 3, 3), 4, 0.3), seeded_randn((1,), 5)
    a = torch.tensor([0.2])
    v = x.permute(0, 2, 3, 1).unsqueeze(1)
    ref = F.conv3d(F.prelu(F.conv3d(v, w1, b1, padding=1), a), w2, b2, padding=1)[:, 0].permute(0, 3, 1, 2)
    kd, mask = ag._band_index(D, "cpu")
    g1 = w1[:, 0].index_select(3, kd).reshape(Cm, 3, 3, D, D) * mask
    g2 = w2[0].index_select(3, kd).reshape(Cm, 3, 3, D, D) * mask
    W1 = g1.permute(3, 0, 4, 1, 2).reshape(D * Cm, D, 3, 3)
    W2 = g2.permute(3, 4, 0, 1, 2).reshape(D, D * Cm, 3, 3)
    hid = F.prelu(F.conv2d(x, W1, b1.repeat(D), padding=1), a)
    out = F.conv2d(hid, W2, b2.expand(D), padding=1)
    assert float((out - ref).abs().max()) < 1e-5


def test_ood_rule():
    from cwfa_b200.training import ood_decision
    nll = [torch.tensor([1.0, 1.5, -2.0]), torch.tensor([0.0, 9.0, 9.0])]
    assert ood_decision(nll, 0, -1.33).tolist() == [False, True, False]       # LL = -NLL below the threshold -> out of distribution
    assert ood_decision(nll, 1, -1.33).tolist() == [False, True, True]

"""ActNorm and AllInOneBlock against golden outputs of the reference's own FrEIA modules
(tests/golden/modules_f4.pt, made by tests/golden/make_golden_f4.py).  fp32 kernels; tolerance rel-L2 5e-5."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, max_abs, rel_l2
from oracle.weights import deterministic_fill, seeded_randn

DEV = "cuda:0"
TOL = 5e-5

AI1_CASES = {
    "hard_softplus_cond": dict(kw=dict(), cond=True),
    "soft_sigmoid_revperm": dict(kw=dict(permute_soft=True, global_affine_type="SIGMOID", reverse_permutation=True, global_affine_init=0.8), cond=True),
    "gin_exp_nocond": dict(kw=dict(gin_block=True, global_affine_type="EXP"), cond=False),
    "householder2": dict(kw=dict(learned_householder_permutation=2, affine_clamping=1.5), cond=True),
}


@pytest.fixture(scope="module")
def fx():
    return torch.load(os.path.join(GOLDEN, "modules_f4.pt"), weights_only=False)


def _ai1(name, fx):
    import cwfa_b200.modules as Fm
    from cwfa_b200 import networks
    networks.networks_n_chans = 64
    spec = AI1_CASES[name]
    ch, H, W = 6, 12, 16
    torch.manual_seed(5); np.random.seed(5)
    m = Fm.AllInOneBlock([(ch, H, W)], dims_c=[(ch, H, W)] * int(spec["cond"]), subnet_constructor=networks.wavelet_flow_subnetwork2D,
                         **spec["kw"]).eval()
    sd = m.state_dict()
    sd.update(deterministic_fill({k: v for k, v in sd.items() if k.startswith("subnet.")}, 400))
    state = fx[f"ai1/{name}/state"]
    assert set(state) == {k for k in sd if not k.startswith("subnet.")}              # same non-subnet keys as the reference
    assert all(tuple(state[k].shape) == tuple(sd[k].shape) for k in state)
    sd.update({k: v.clone() for k, v in state.items()})
    m.load_state_dict(sd)
    return m, spec


@pytest.mark.parametrize("name", list(AI1_CASES))
def test_all_in_one_block_state_dict_matches_reference(fx, name):
    _ai1(name, fx)                                                                    # CPU: keys / shapes only


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(AI1_CASES))
def test_all_in_one_block_vs_reference(fx, name):
    m, spec = _ai1(name, fx)
    m = m.to(DEV)
    x = seeded_randn((2, 6, 12, 16), 81).to(DEV)
    c = [seeded_randn((2, 6, 12, 16), 82).to(DEV)] if spec["cond"] else []
    with torch.no_grad():
        (y,), j = m((x,), c=c)
        (xr,), jr = m((x,), c=c, rev=True)
        (xb,), jb = m((y,), c=c, rev=True)                                            # round trip
    assert rel_l2(y, fx[f"ai1/{name}/fwd"]) < TOL and rel_l2(xr, fx[f"ai1/{name}/rev"]) < TOL
    for got, ref in ((j, fx[f"ai1/{name}/fwd_jac"]), (jr, fx[f"ai1/{name}/rev_jac"])):
        assert max_abs(got, ref) < 1e-4 * max(1.0, float(ref.abs().max())) + 2e-3
    assert rel_l2(xb, x) < 1e-4 and max_abs(j + jb, torch.zeros_like(j)) < 2e-2


@pytest.mark.gpu
def test_actnorm_vs_reference(fx):
    import cwfa_b200.modules as Fm
    x = (seeded_randn((3, 6, 8, 10), 80) * 1.7 + 0.4).to(DEV)
    m = Fm.ActNorm([(6, 8, 10)]).to(DEV)
    with torch.no_grad():
        (y,), j = m((x,))                                                             # initialises from this first batch
        (xr,), jr = m((y,), rev=True)
    assert rel_l2(m.scale, fx["actnorm/scale"]) < 1e-5 and max_abs(m.bias, fx["actnorm/bias"]) < 1e-5
    assert rel_l2(y, fx["actnorm/fwd"]) < TOL and rel_l2(xr, fx["actnorm/rev"]) < TOL
    assert max_abs(j, fx["actnorm/jac"]) < 1e-3 and max_abs(jr, fx["actnorm/rev_jac"]) < 1e-3
    # a loaded state_dict must not be re-initialised by the next batch (invertible_resnet.py:46-52)
    m2 = Fm.ActNorm([(6, 8, 10)]).to(DEV)
    m2.load_state_dict(m.state_dict())
    with torch.no_grad():
        (y2,), _ = m2((x * 3.0,))
    assert not m2.init_on_next_batch and rel_l2(y2, (x * 3.0) * m.scale.exp() + m.bias) < 1e-5


# ---- training: ActNorm / AllInOneBlock are differentiable end to end (scale_shift + 1x1 mixing adjoint kernels) ----
def _nll_like(y, j, r):
    """A loss that touches both outputs the way the NLL does: <y, r> + 0.5 ||y||^2 / numel - mean(j) / numel."""
    return (y * r).sum() + 0.5 * (y ** 2).sum() / y.numel() - j.mean() / y[0].numel()


@pytest.mark.gpu
def test_actnorm_gradients_vs_oracle_autograd(fx):
    """d loss / d (x, scale, bias) of ActNorm in both directions: this repo's adjoint kernels vs torch CPU autograd through the
    oracle's restatement (oracle/cwfa_oracle.py:actnorm, pinned to the reference golden above)."""
    import cwfa_b200.modules as Fm
    from oracle import cwfa_oracle as O
    x0 = seeded_randn((3, 6, 8, 10), 80) * 1.7 + 0.4
    r = seeded_randn((3, 6, 8, 10), 85)
    for rev in (False, True):
        m = Fm.ActNorm([(6, 8, 10)]).to(DEV)
        with torch.no_grad():
            m((x0.to(DEV),))                                    # data-dependent init
            m.scale += 0.3 * seeded_randn(tuple(m.scale.shape), 86).to(DEV)
            m.bias += 0.2 * seeded_randn(tuple(m.bias.shape), 87).to(DEV)
        x = x0.to(DEV).requires_grad_(True)
        (y,), j = m((x,), rev=rev)
        _nll_like(y, j, r.to(DEV)).backward()
        sc = m.scale.detach().cpu().clone().requires_grad_(True)
        bi = m.bias.detach().cpu().clone().requires_grad_(True)
        xc = x0.clone().requires_grad_(True)
        yo, jo = O.actnorm(xc, sc, bi, rev=rev)
        _nll_like(yo, jo, r).backward()
        assert rel_l2(y, yo) < TOL
        for name, got, ref in (("x", x.grad, xc.grad), ("scale", m.scale.grad, sc.grad), ("bias", m.bias.grad, bi.grad)):
            assert got is not None, f"ActNorm: no gradient reached {name}"
            assert rel_l2(got, ref) < 2e-4, (rev, name, rel_l2(got, ref))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["hard_softplus_cond", "householder2", "soft_sigmoid_revperm"])
def test_all_in_one_block_gradients_vs_oracle_autograd(fx, name):
    """global_scale, global_offset, the Householder vectors and the sub-network all receive the gradients torch CPU autograd
    computes through the oracle's restatement of AllInOneBlock (pinned to the reference goldens in test_oracle_golden.py)."""
    from oracle import cwfa_oracle as O
    m, spec = _ai1(name, fx)
    m = m.to(DEV)
    x0 = seeded_randn((2, 6, 12, 16), 81)
    c0 = seeded_randn((2, 6, 12, 16), 82)
    r = seeded_randn((2, 6, 12, 16), 88)
    x = x0.to(DEV).requires_grad_(True)
    (y,), j = m((x,), c=[c0.to(DEV)] if spec["cond"] else [])
    _nll_like(y, j, r.to(DEV)).backward()
    sd = {"m." + k: (v.detach().cpu().clone().requires_grad_(True) if v.dtype.is_floating_point else v.detach().cpu())
          for k, v in m.state_dict(keep_vars=True).items()}
    xc = x0.clone().requires_grad_(True)
    kw = spec["kw"]
    yo, jo = O.all_in_one_block(sd, "m.", xc, [c0] if spec["cond"] else [], False, clamp=kw.get("affine_clamping", 2.0),
                                gin=kw.get("gin_block", False), global_affine_type=kw.get("global_affine_type", "SOFTPLUS"),
                                reverse_permutation=kw.get("reverse_permutation", False),
                                householder=kw.get("learned_householder_permutation", 0))
    _nll_like(yo, jo, r).backward()
    assert rel_l2(y, yo) < TOL
    assert rel_l2(x.grad, xc.grad) < 5e-4
    checked = 0
    for k, p in m.named_parameters():
        if not p.requires_grad:
            continue                                             # frozen permutation matrices (w_perm / w_perm_inv / w_0)
        ref = sd["m." + k].grad if sd["m." + k].dtype.is_floating_point else None
        if ref is None or float(ref.abs().max()) == 0.0:
            continue                                             # unused checkpoint-compat weights / frozen permutation matrices
        assert p.grad is not None, f"AllInOneBlock[{name}]: no gradient reached {k}"
        assert rel_l2(p.grad, ref) < 2e-3, (k, rel_l2(p.grad, ref))
        checked += 1
    must = {"global_scale", "global_offset"} | ({"vk_householder"} if kw.get("learned_householder_permutation") else set())
    assert must <= {k for k, p in m.named_parameters() if p.grad is not None and float(p.grad.abs().max()) > 0}
    assert checked >= 10

"""GPU: whole-path parity on BASELINE.json configs[0] (tiny CWFA: D=16, S=64, 3 steps) against the golden
outputs of the unmodified reference, through the reference-facing module API (fp32 kernels)."""
import pytest
import torch

from conftest import max_abs, rel_l2
from helpers import build_tiny_model, tiny_inputs
from oracle.weights import seeded_randn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4        # rel-L2, fp32 path (stated tolerance; SURVEY.md section 7 'Precision')


@pytest.fixture(scope="module")
def model(golden_tiny):
    m = build_tiny_model(golden_tiny, DEV)
    m._lrnn_snapshot = {k: v.detach().clone() for k, v in m.cond_nets[-1].state_dict().items()}
    return m


@pytest.mark.parametrize("bn_mode", ["batch", "running"])
@pytest.mark.parametrize("use_mv", [False, True])
def test_inverse_reconstruction_vs_reference(golden_tiny, model, bn_mode, use_mv):
    views, mean_vols = tiny_inputs(golden_tiny)
    L = model.n_levels
    mv = [t.to(DEV) for t in mean_vols[:L]] + ([mean_vols[L].to(DEV)] if use_mv else [None])
    # a .train()-mode forward updates the BatchNorm running statistics (as in the reference): start every case from the
    # fixture's weights, like the golden generator does (tests/golden/make_golden.py: fill(enc, 300) per case)
    model.cond_nets[-1].load_state_dict(model._lrnn_snapshot)
    model.cond_nets[-1].train(bn_mode == "batch")
    outs, jacs = model.reconstruct(views.to(DEV), mv, return_all=True)
    model.cond_nets[-1].train()
    tag = f"recon/{bn_mode}/{'mv' if use_mv else 'nomv'}"
    errs = {"lrnn": (rel_l2(outs[L], golden_tiny[f"{tag}/lrnn"]), max_abs(outs[L], golden_tiny[f"{tag}/lrnn"]))}
    for n in range(L):
        errs[n] = (rel_l2(outs[n], golden_tiny[f"{tag}/vol{n}"]), max_abs(outs[n], golden_tiny[f"{tag}/vol{n}"]))
    print(tag, {k: (f"{a:.2e}", f"{b:.2e}") for k, (a, b) in errs.items()})
    assert all(a < TOL for a, _ in errs.values()), errs
    for n in range(L):
        ref = float(golden_tiny[f"{tag}/jac{n}"][0])
        assert abs(float(jacs[n][0]) - ref) < 1e-4 * max(1.0, abs(ref)) + 1e-2


def test_forward_pyramid_vs_reference(golden_tiny, model):
    cfg = golden_tiny["config"]
    B, D, S = 2, cfg["D"], cfg["S"]
    x = seeded_randn((B, D, S, S), 2).to(DEV)
    vB = seeded_randn((B, 29, S, S), 3).to(DEV)
    _, mean_vols = tiny_inputs(golden_tiny)
    res = model.forward_nll(x, vB, [mv.repeat(B, 1, 1, 1).to(DEV) for mv in mean_vols[:model.n_levels]])
    for n, r in enumerate(res):
        assert rel_l2(r["z"], golden_tiny[f"fwd/z{n}"]) < TOL
        assert rel_l2(r["lo"], golden_tiny[f"fwd/lo{n}"]) < 1e-6
        assert rel_l2(r["logdet"], golden_tiny[f"fwd/jac{n}"]) < 1e-4
        assert rel_l2(r["nll_ref"], golden_tiny[f"fwd/nll_ref{n}"]) < 1e-4


def test_round_trip_and_logdet_sign(golden_tiny, model):
    cfg = golden_tiny["config"]
    B, D, S = 2, cfg["D"], cfg["S"]
    x = seeded_randn((B, D, S, S), 12).to(DEV)
    vB = seeded_randn((B, 29, S, S), 13).to(DEV)
    _, mean_vols = tiny_inputs(golden_tiny)
    c0 = model.cond_nets[0](vB)[-1]
    c = [c0, mean_vols[0].repeat(B, 1, 1, 1).to(DEV)]
    (z, lo), j = model.conv_inn[0](x, c=c)
    xr, jr = model.conv_inn[0]([z, lo], c=c, rev=True)
    assert rel_l2(xr, x) < 1e-5
    assert max_abs(j, -jr) < 1e-3 * float(j.abs().max())


def test_frames_are_independent(golden_tiny, model):
    """Sharding premise: a batch of frames == the frames one by one, bit for bit
    (flow levels; no cross-frame arithmetic)."""
    cfg = golden_tiny["config"]
    D, S = cfg["D"], cfg["S"]
    vB = seeded_randn((3, 29, S, S), 31).to(DEV)
    _, mean_vols = tiny_inputs(golden_tiny)
    lo = seeded_randn((3, D // 2, S, S), 32).to(DEV)
    mv = mean_vols[0].repeat(3, 1, 1, 1).to(DEV)
    z = torch.zeros_like(lo)
    full, jf = model.conv_inn[0]([z, lo], c=[model.cond_nets[0](vB)[-1], mv], rev=True)
    for b in range(3):
        one, j1 = model.conv_inn[0]([z[b:b + 1], lo[b:b + 1]], c=[model.cond_nets[0](vB[b:b + 1])[-1], mv[b:b + 1]], rev=True)
        assert torch.equal(one[0], full[b]) and torch.equal(j1[0], jf[b])

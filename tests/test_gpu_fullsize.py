"""GPU: BASELINE.json configs[1] / configs[2] at FULL size (512x512x96, 5 steps): the tensor-core engine against the
CPU oracle on the same random-init weights, plus size-independent properties (round trip, log-det sign)."""
import pytest
import torch

from conftest import max_abs, rel_l2
from oracle import cwfa_oracle as O
from oracle.weights import seeded_randn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def full():
    import cwfa_b200
    model = cwfa_b200.CWFAModel(seed=0)            # defaults = main.py defaults: 96 x 512 x 512, 5 steps, CAT, 4 blocks
    om = model.export_for_oracle()
    views = seeded_randn((1, 29, 512, 512), 1)
    mvs = [seeded_randn((1, 96 // 2 ** (n + 1), 512, 512), 10 + n, 0.1) for n in range(4)] + [seeded_randn((1, 6, 512, 512), 14, 0.1)]
    return model.to(DEV), om, views, mvs


_REF = {}


def _oracle_inverse(om, views, mvs):
    """The fp32 CPU oracle of the full frame (tens of seconds): computed once per session."""
    if "inv" not in _REF:
        _REF["inv"] = O.reconstruct(om, views, mvs, bn_mode="batch", return_all=True)
    return _REF["inv"]


# tolerances on the per-level rel-L2 vs the fp32 oracle: half-precision operand rounding accumulated through ~40 convs
@pytest.mark.parametrize("kind,tol", [("bf16", 2e-2), ("fp16", 4e-3)])
def test_full_inverse_engine_vs_cpu_oracle(full, kind, tol):
    from cwfa_b200.engine import CWFAEngine
    model, om, views, mvs = full
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ref, ref_j = _oracle_inverse(om, views, mvs)
    eng = CWFAEngine(model, kind)
    outs, jacs = eng.reconstruct(views.to(DEV), [m.to(DEV) for m in mvs], return_all=True)
    rep = {n: (rel_l2(outs[n], ref[n]), max_abs(outs[n], ref[n])) for n in sorted(ref)}
    print(f"full 512x512x96 inverse, {kind} engine vs fp32 CPU oracle (rel-L2, max-abs) per level:",
          {k: (f"{a:.2e}", f"{b:.2e}") for k, (a, b) in rep.items()})
    assert all(a < tol for a, _ in rep.values()), rep
    for n in ref_j:
        r = float(ref_j[n][0])
        assert abs(float(jacs[n][0]) - r) < tol * max(1.0, abs(r)), (n, float(jacs[n][0]), r)
    assert tuple(outs[0].shape) == (1, 96, 512, 512)
    g = eng.reconstruct_graphed(views.to(DEV), [m.to(DEV) for m in mvs])
    assert torch.equal(g, outs[0]), "CUDA-graph replay (parallel branches) must equal the eager engine bit for bit"


def test_full_forward_nll_and_round_trip(full):
    from cwfa_b200.engine import CWFAEngine
    model, om, views, mvs = full
    B = 2
    x = seeded_randn((B, 96, 512, 512), 2).to(DEV)
    vB = seeded_randn((B, 29, 512, 512), 3).to(DEV)
    mv = [m.repeat(B, 1, 1, 1).to(DEV) for m in mvs[:4]]
    eng = CWFAEngine(model, "bf16")
    res = eng.forward_nll(x, vB, mv)
    # level 0 through the fp32 module path: forward, then the exact inverse (round trip), and the sign convention
    c0 = model.cond_nets[0](vB)[-1]
    (z, lo), j = model.conv_inn[0](x, c=[c0, mv[0]])
    xr, jr = model.conv_inn[0]([z, lo], c=[c0, mv[0]], rev=True)
    assert rel_l2(xr, x) < 1e-5
    assert max_abs(j, -jr) < 1e-4 * float(j.abs().max())
    e = (rel_l2(res[0]["z"], z), rel_l2(res[0]["logdet"], j), rel_l2(res[0]["lo"], lo))
    print("full forward level 0, bf16 engine vs fp32 module path: z", f"{e[0]:.2e}", "logdet", f"{e[1]:.2e}")
    assert e[0] < 2e-2 and e[1] < 2e-2 and e[2] < 1e-6
    # Haar pyramid is orthonormal: energy is preserved level by level (size-independent property)
    en = float((x.double() ** 2).sum())
    assert abs(float((lo.double() ** 2).sum() + ((O.haar1d(x.cpu())[0][:, 48:]).double() ** 2).sum()) - en) < 1e-6 * en
    assert all(torch.isfinite(r["nll_per_sample"]).all() for r in res)

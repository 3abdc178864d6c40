"""GPU: a RAGGED configuration end to end -- 72 x 72 pixels (4.5 tiles of 16, 9 M-blocks of 8 columns: every tensor-core kernel
runs partial tiles), 24 depths (12 / 6 / 6 channels: F8 groups and 16-channel blocks with padding), 3 steps, batch 2 (BatchNorm
batch statistics over two frames) -- against the CPU oracle on the same seeded weights and inputs: fp32 module path, tensor-core
module path, engine (eager and CUDA-graph replay), forward pyramid with per-level log-det / sum z^2."""
import pytest
import torch

from conftest import max_abs, rel_l2
from oracle import cwfa_oracle as O
from oracle.weights import deterministic_fill, seeded_randn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
D, S, L, B = 24, 72, 3, 2


@pytest.fixture(scope="module")
def ragged():
    import cwfa_b200
    model = cwfa_b200.CWFAModel(n_depths=D, volume_side_size=S, INN_max_down_steps=L, seed=0)
    for n in range(model.n_levels):
        sd = deterministic_fill(model.conv_inn[n].state_dict(), 400 + n)
        sd.update({k: v for k, v in model.conv_inn[n].state_dict().items() if "perm" in k})
        model.conv_inn[n].load_state_dict(sd)
        model.cond_nets[n].load_state_dict(deterministic_fill(model.cond_nets[n].state_dict(), 500 + n))
    model.cond_nets[-1].load_state_dict(deterministic_fill(model.cond_nets[-1].state_dict(), 600))
    om = model.export_for_oracle()
    views = seeded_randn((B, 29, S, S), 41)
    mvs = [seeded_randn((B, D // 2 ** (n + 1), S, S), 50 + n, 0.1) for n in range(L - 1)] + [seeded_randn((B, D // 2 ** (L - 1), S, S), 59, 0.1)]
    return model.to(DEV), om, views, mvs


def test_ragged_inverse_all_paths_vs_cpu_oracle(ragged):
    import cwfa_b200
    from cwfa_b200.engine import CWFAEngine
    model, om, views, mvs = ragged
    ref, ref_j = O.reconstruct(om, views, mvs, bn_mode="batch", return_all=True)
    vd, md = views.to(DEV), [m.to(DEV) for m in mvs]
    out32 = model.reconstruct(vd, md)
    assert tuple(out32.shape) == (B, D, S, S)
    assert rel_l2(out32, ref[0]) < 1e-4, rel_l2(out32, ref[0])
    for kind, tol in (("bf16", 2e-2), ("fp16", 3e-3)):
        eng = CWFAEngine(model, kind)
        outs, jacs = eng.reconstruct(vd, md, return_all=True)
        rep = {n: rel_l2(outs[n], ref[n]) for n in sorted(ref)}
        print(f"ragged {S}x{S}x{D} batch {B}, {kind} engine vs fp32 CPU oracle, rel-L2 per level:", {k: f"{v:.2e}" for k, v in rep.items()})
        assert all(v < tol for v in rep.values()), rep
        for n in ref_j:
            for b in range(B):
                r = float(ref_j[n][b])
                assert abs(float(jacs[n][b]) - r) < tol * max(1.0, abs(r)), (kind, n, b, float(jacs[n][b]), r)
        assert torch.equal(eng.reconstruct_graphed(vd, md), outs[0]), "graph replay must equal the eager engine bit for bit"
        with cwfa_b200.inference_precision(kind):
            out_api = model.reconstruct(vd, md)
        assert rel_l2(out_api, ref[0]) < tol, (kind, rel_l2(out_api, ref[0]))


def test_ragged_forward_pyramid_vs_cpu_oracle(ragged):
    from cwfa_b200.engine import CWFAEngine
    model, om, views, mvs = ragged
    x = seeded_randn((B, D, S, S), 42)
    ref = O.forward_nll(om, x, views, mvs[:L - 1])
    vd, md = views.to(DEV), [m.to(DEV) for m in mvs[:L - 1]]
    res32 = model.forward_nll(x.to(DEV), vd, md)
    for n, (r, g) in enumerate(zip(ref, res32)):
        assert rel_l2(g["z"], r["z"]) < 1e-4 and rel_l2(g["lo"], r["lo"]) < 1e-5, n
        assert max_abs(g["logdet"], r["logdet"]) < 1e-4 * max(1.0, float(r["logdet"].abs().max())), n
    for kind, tol in (("bf16", 2e-2), ("fp16", 3e-3)):
        res = CWFAEngine(model, kind).forward_nll(x.to(DEV), vd, md)
        for n, (r, g) in enumerate(zip(ref, res)):
            e = (rel_l2(g["z"], r["z"]), rel_l2(g["lo"], r["lo"]), rel_l2(g["logdet"], r["logdet"]), rel_l2(g["sumsq"], r["sumsq"]))
            print(f"ragged forward level {n}, {kind}: z {e[0]:.2e} lo {e[1]:.2e} logdet {e[2]:.2e} sumsq {e[3]:.2e}")
            assert e[0] < tol and e[1] < 1e-5 and e[2] < tol and e[3] < 2 * tol, (kind, n, e)

"""GPU: the tensor-core engine (bf16 / fp16 operands, fp32 accumulation) against the golden outputs of the
unmodified reference on BASELINE.json configs[0], with the stated half-precision tolerances, reported per
level (max-abs and rel-L2)."""
import pytest
import torch

from conftest import max_abs, rel_l2
from helpers import build_tiny_model, tiny_inputs
from oracle.weights import seeded_randn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = {"bf16": 2e-2, "fp16": 3e-3}          # rel-L2 on volumes; log-det: same factor relative to |J|


@pytest.fixture(scope="module")
def model(golden_tiny):
    return build_tiny_model(golden_tiny, DEV)


@pytest.mark.parametrize("kind", ["bf16", "fp16"])
@pytest.mark.parametrize("bn_mode", ["batch", "running"])
@pytest.mark.parametrize("use_mv", [False, True])
def test_engine_inverse_vs_reference(golden_tiny, model, kind, bn_mode, use_mv):
    from cwfa_b200.engine import CWFAEngine
    views, mean_vols = tiny_inputs(golden_tiny)
    L = model.n_levels
    mv = [t.to(DEV) for t in mean_vols[:L]] + ([mean_vols[L].to(DEV)] if use_mv else [None])
    model.cond_nets[-1].train(bn_mode == "batch")
    eng = CWFAEngine(model, kind)
    outs, jacs = eng.reconstruct(views.to(DEV), mv, return_all=True)
    model.cond_nets[-1].train()
    tag = f"recon/{bn_mode}/{'mv' if use_mv else 'nomv'}"
    rep = {"lrnn": (rel_l2(outs[L], golden_tiny[f"{tag}/lrnn"]), max_abs(outs[L], golden_tiny[f"{tag}/lrnn"]))}
    for n in range(L):
        rep[n] = (rel_l2(outs[n], golden_tiny[f"{tag}/vol{n}"]), max_abs(outs[n], golden_tiny[f"{tag}/vol{n}"]))
    print(kind, tag, {k: (f"{a:.2e}", f"{b:.2e}") for k, (a, b) in rep.items()})
    assert all(a < TOL[kind] for a, _ in rep.values()), rep
    for n in range(L):
        ref = float(golden_tiny[f"{tag}/jac{n}"][0])
        assert abs(float(jacs[n][0]) - ref) < TOL[kind] * max(1.0, abs(ref)), (n, float(jacs[n][0]), ref)


@pytest.mark.parametrize("kind", ["bf16", "fp16"])
def test_engine_forward_nll_vs_reference(golden_tiny, model, kind):
    from cwfa_b200.engine import CWFAEngine
    cfg = golden_tiny["config"]
    B, D, S = 2, cfg["D"], cfg["S"]
    x = seeded_randn((B, D, S, S), 2).to(DEV)
    vB = seeded_randn((B, 29, S, S), 3).to(DEV)
    _, mean_vols = tiny_inputs(golden_tiny)
    eng = CWFAEngine(model, kind)
    res = eng.forward_nll(x, vB, [mv.repeat(B, 1, 1, 1).to(DEV) for mv in mean_vols[:model.n_levels]])
    for n, r in enumerate(res):
        e = (rel_l2(r["z"], golden_tiny[f"fwd/z{n}"]), rel_l2(r["logdet"], golden_tiny[f"fwd/jac{n}"]),
             rel_l2(r["nll_ref"], golden_tiny[f"fwd/nll_ref{n}"]))
        print(kind, "fwd level", n, [f"{v:.2e}" for v in e])
        assert e[0] < TOL[kind] and e[1] < TOL[kind] and e[2] < TOL[kind]
        assert rel_l2(r["lo"], golden_tiny[f"fwd/lo{n}"]) < 1e-6


def test_engine_matches_fp32_module_path_and_graph_replay(golden_tiny, model):
    from cwfa_b200.engine import CWFAEngine
    views, mean_vols = tiny_inputs(golden_tiny)
    mv = [t.to(DEV) for t in mean_vols]
    ref = model.reconstruct(views.to(DEV), mv)
    eng = CWFAEngine(model, "bf16")
    a = eng.reconstruct(views.to(DEV), mv)
    assert rel_l2(a, ref) < TOL["bf16"]
    g1 = eng.reconstruct_graphed(views.to(DEV), mv).clone()
    g2 = eng.reconstruct_graphed(views.to(DEV), mv).clone()
    assert torch.equal(g1, g2), "graph replay must be bit-reproducible"
    assert torch.equal(g1, a), "graph replay must equal the eager engine bit for bit"


def test_unet_pieces_c8(golden_tiny):
    """convT scatter epilogue, C8 BatchNorm (+pool) against torch CPU ops."""
    import torch.nn.functional as F
    from cwfa_b200 import tc
    x = seeded_randn((2, 32, 8, 16), 1).bfloat16().float()
    w = (seeded_randn((32, 48, 2, 2), 2) * 0.2).bfloat16().float()
    b = seeded_randn((48,), 3)
    skip = seeded_randn((2, 48, 16, 32), 4).bfloat16().float()
    ref = F.conv_transpose2d(x, w, b, stride=2) + skip
    pc = tc.PackedConv(w.to(DEV), b.to(DEV), "bf16", transposed=True)
    y = tc.from_c8(tc.conv_transpose_tc(tc.to_c8(x.to(DEV)), pc, tc.to_c8(skip.to(DEV))))
    assert rel_l2(y, ref) < 4e-3
    xb = seeded_randn((2, 256, 8, 12), 5).bfloat16().float()
    g, be = seeded_randn((256,), 6) * 0.2 + 1.0, seeded_randn((256,), 7)
    refb = F.batch_norm(xb, None, None, g, be, training=True)
    yb, yp = tc.batchnorm_c8(tc.to_c8(xb.to(DEV)), g.to(DEV), be.to(DEV), None, None, batch_stats=True, pool=True)
    assert rel_l2(tc.from_c8(yb), refb) < 4e-3
    assert rel_l2(tc.from_c8(yp), F.max_pool2d(refb, 2)) < 4e-3
    assert torch.equal(tc.from_c8(yp).cpu(), F.max_pool2d(tc.from_c8(yb).cpu(), 2))
    rm, rv = seeded_randn((256,), 8) * 0.1, seeded_randn((256,), 9).abs() + 0.5
    refr = F.batch_norm(xb, rm, rv, g, be, training=False)
    yr = tc.batchnorm_c8(tc.to_c8(xb.to(DEV)), g.to(DEV), be.to(DEV), rm.to(DEV), rv.to(DEV), batch_stats=False)
    assert rel_l2(tc.from_c8(yr), refr) < 4e-3


def test_streaming_host_api_matches_engine(golden_tiny, model):
    from cwfa_b200.engine import CWFAEngine, StreamingReconstructor
    views, mean_vols = tiny_inputs(golden_tiny)
    mv = [t.to(DEV) for t in mean_vols]
    eng = CWFAEngine(model, "bf16")
    frames = [seeded_randn(tuple(views.shape), 50 + i).pin_memory() for i in range(5)]
    outs = [torch.empty((1, golden_tiny["config"]["D"], views.shape[2], views.shape[3]), dtype=torch.float32, pin_memory=True)
            for _ in range(5)]
    StreamingReconstructor(eng, tuple(views.shape), mv, depth=2).run(frames, outs)
    for f, o in zip(frames, outs):
        ref = eng.reconstruct(f.to(DEV), mv)
        assert torch.equal(o, ref.cpu()), "streamed frames must equal the one-by-one result bit for bit"
    one = eng.reconstruct_host(frames[2], mv)
    assert torch.equal(one, outs[2])


def test_engine_inverse_with_latent_samples(golden_tiny, model):
    """Temperature > 0: the engine's inverse with explicit latent samples z (CWFA.py:906-912) against the fp32 module
    path of the same model; z enters through the final PermuteRandom^-1 and the gather of the last coupling."""
    from cwfa_b200.engine import CWFAEngine
    views, mvs = tiny_inputs(golden_tiny)
    views, mvs = views.to(DEV), [m.to(DEV) for m in mvs]
    zs = [0.7 * seeded_randn((1,) + tuple(model.conv_inn[n].global_out_shapes[0]), 70 + n).to(DEV) for n in range(model.n_levels)]
    ref, ref_j = model.reconstruct(views, mvs, zs=zs, return_all=True)
    out0 = model.reconstruct(views, mvs)
    eng = CWFAEngine(model, "fp16")
    outs, jacs = eng.reconstruct(views, mvs, return_all=True, zs=zs)
    for n in ref:
        assert rel_l2(outs[n], ref[n]) < 5e-3, (n, rel_l2(outs[n], ref[n]))
    for n in ref_j:
        assert abs(float(jacs[n][0]) - float(ref_j[n][0])) < 5e-3 * max(1.0, abs(float(ref_j[n][0])))
    assert rel_l2(ref[0], out0) > 1e-2          # the samples matter

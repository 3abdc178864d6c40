"""GPU: the tensor-core engine (bf16 / fp16 operands, fp32 accumulation) against the golden outputs of the
unmodified reference on BASELINE.json configs[0], with the stated half-precision tolerances, reported per
level (max-abs and rel-L2)."""
import pytest
import torch

from conftest import max_abs, rel_l2
from helpers import build_tiny_model, tiny_inputs
from oracle import cwfa_oracle as O
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
    # fp16 host volumes (half the D2H bytes; the reference's own GPU output dtype under autocast, CWFA.py:845): exactly the
    # round-to-nearest fp16 of the fp32 result (one cast kernel on the device, cwfa_cast_f32_f16)
    outs16 = [torch.empty((1, golden_tiny["config"]["D"], views.shape[2], views.shape[3]), dtype=torch.float16, pin_memory=True) for _ in range(5)]
    StreamingReconstructor(eng, tuple(views.shape), mv, depth=2, out_dtype=torch.float16).run(frames, outs16)
    for o16, o in zip(outs16, outs):
        assert torch.equal(o16, o.half())
    from cwfa_b200 import ops
    x = seeded_randn((3, 5, 7, 11), 9).to(DEV) * 100.0                  # odd element count: vector body + scalar tail
    assert torch.equal(ops.cast_f16(x), x.half())


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


# ---- the F8 coupling path (csrc/coupling_f8.cu) ------------------------------------------------------------------------
@pytest.mark.parametrize("B,C,H,W", [(1, 16, 64, 64), (2, 12, 8, 12), (1, 96, 32, 32), (3, 4, 6, 6)])
def test_f8_haar_and_converters_vs_oracle(B, C, H, W):
    """Depth-wise Haar + Split with the detail half in the F8 layout, and the NCHW <-> F8 converters with a channel map, against
    the oracle's Haar (INN_utils.py:142-161) -- bit-level: same arithmetic, other layout."""
    from cwfa_b200 import tc
    x = seeded_randn((B, C, H, W), 70)
    y, _ = O.haar1d(x)
    h = C // 2
    lo, hi8 = tc.haar1d_split_f8(x.to(DEV))
    assert max_abs(lo, y[:, :h]) < 1e-6 and max_abs(tc.from_f8(hi8, h), y[:, h:]) < 1e-6
    if h % 8:
        assert float(hi8.reshape(B, -1, H * W, 8)[:, -1, :, h % 8:].abs().max()) == 0.0        # channel padding is zero
    assert max_abs(tc.haar1d_merge_f8(lo, hi8), x) < 1e-6
    perm = torch.from_numpy(__import__("numpy").random.RandomState(3).permutation(h)).to(torch.int32).to(DEV)
    z = seeded_randn((B, h, H, W), 71).to(DEV)
    z8 = tc.to_f8(z, perm)                                                                   # slot j = channel perm[j]
    assert torch.equal(tc.from_f8(z8, h), z[:, perm.long()])
    inv = torch.empty_like(perm)
    inv[perm.long()] = torch.arange(h, device=DEV, dtype=torch.int32)
    assert torch.equal(tc.from_f8(z8, h, inv), z)                                            # channel c = slot inv[c]


@pytest.mark.parametrize("kind", ["bf16", "fp16"])
def test_engine_f8_path_equals_nchw_path(golden_tiny, model, kind):
    """The lean F8 coupling path (channel permutations folded into the packed weights, bias through an MMA, 128-bit I/O) against
    the NCHW coupling path of the same engine: inverse (z = 0 and given z) and forward NLL; differences are fp32 rounding only."""
    from cwfa_b200.engine import CWFAEngine
    views, mean_vols = tiny_inputs(golden_tiny)
    views, mv = views.to(DEV), [t.to(DEV) for t in mean_vols]
    eng = CWFAEngine(model, kind)
    assert all(lv["f8"] is not None for lv in eng.levels)
    zs = [seeded_randn((1,) + tuple(model.conv_inn[n].global_out_shapes[0]), 130 + n, 0.5).to(DEV) for n in range(model.n_levels)]
    cfg = golden_tiny["config"]
    x = seeded_randn((2, cfg["D"], cfg["S"], cfg["S"]), 2).to(DEV)
    vB = seeded_randn((2, 29, cfg["S"], cfg["S"]), 3).to(DEV)
    mvB = [t.repeat(2, 1, 1, 1) for t in mv[:model.n_levels]]
    res = {}
    for f8 in (True, False):
        eng.use_f8 = f8
        res[f8] = (eng.reconstruct(views, mv, return_all=True), eng.reconstruct(views, mv, zs=zs, return_all=True), eng.forward_nll(x, vB, mvB))
    for a, b in ((res[True][0], res[False][0]), (res[True][1], res[False][1])):
        for n in a[0]:
            assert rel_l2(a[0][n], b[0][n]) < 2e-5, (n, rel_l2(a[0][n], b[0][n]))
        for n in a[1]:
            assert abs(float(a[1][n][0] - b[1][n][0])) < 1e-4 * max(1.0, abs(float(b[1][n][0])))
    for ra, rb in zip(res[True][2], res[False][2]):
        assert rel_l2(ra["z"], rb["z"]) < 2e-5 and rel_l2(ra["logdet"], rb["logdet"]) < 1e-4 and rel_l2(ra["sumsq"], rb["sumsq"]) < 1e-4
        assert torch.equal(ra["lo"], rb["lo"])


@pytest.mark.parametrize("N,Cin,Cout,H,W", [(1, 256, 256, 40, 24), (2, 128, 512, 16, 16), (1, 512, 1024, 19, 13)])
def test_conv_with_fused_batchnorm_statistics(N, Cin, Cout, H, W):
    """conv -> PReLU -> BatchNorm (unet.py:99-107) with the batch statistics accumulated in the conv's epilogue
    (cwfa_conv_tc_bn + cwfa_bn_partial_finalize) against torch CPU ops, ragged tiles included; the statistics are
    bit-reproducible (every partial is written once, fixed-order two-stage sum) and the conv output itself is unchanged."""
    import torch.nn.functional as F
    from cwfa_b200 import ops, tc
    x = seeded_randn((N, Cin, H, W), 1).bfloat16().float()
    w = (seeded_randn((Cout, Cin, 3, 3), 2) * (2.0 / (9 * Cin)) ** 0.5).bfloat16().float()
    b = seeded_randn((Cout,), 3) * 0.1
    slope = torch.tensor([0.23])
    g, be = seeded_randn((Cout,), 6) * 0.2 + 1.0, seeded_randn((Cout,), 7)
    act = F.prelu(F.conv2d(x, w, b, padding=1), slope)
    ref = F.batch_norm(act, None, None, g, be, training=True)
    pc = tc.PackedConv(w.to(DEV), b.to(DEV), "bf16")
    x8 = tc.to_c8(x.to(DEV))
    y8, part, mb = tc.conv_tc_bn_stats(x8, pc, act=ops.ACT_PRELU, slope=slope.to(DEV))
    assert part is not None, "these shapes must take the wide kernel (fused statistics)"
    assert torch.equal(y8.data, tc.conv_tc(x8, pc, act=ops.ACT_PRELU, slope=slope.to(DEV), mb=mb).data)
    out = tc.batchnorm_c8(y8, g.to(DEV), be.to(DEV), None, None, batch_stats=True, partial=(part, mb))
    assert rel_l2(tc.from_c8(out), ref) < 5e-3
    # against the separate statistics pass over the (half-rounded) stored tensor: same normalisation up to that rounding
    sep = tc.batchnorm_c8(y8, g.to(DEV), be.to(DEV), None, None, batch_stats=True)
    assert rel_l2(tc.from_c8(out), tc.from_c8(sep)) < 3e-3
    y8b, part2, _ = tc.conv_tc_bn_stats(x8, pc, act=ops.ACT_PRELU, slope=slope.to(DEV))
    out2 = tc.batchnorm_c8(y8b, g.to(DEV), be.to(DEV), None, None, batch_stats=True, partial=(part2, mb))
    assert torch.equal(out.data, out2.data)

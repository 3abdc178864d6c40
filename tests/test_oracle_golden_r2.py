"""CPU: the oracle against the ROUND-2 goldens of the unmodified reference (tests/golden/make_golden_r2.py): SequenceINN, the
LRNN's mean-volume convention (CWFA.py:882), disable_low_res_input=1, TANH / SIGMOID / callable clamps, the multi-sample path,
a checkpoint file WRITTEN BY THE REFERENCE, and the probe of the reference at the FULL 96 x 512 x 512 config."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, max_abs, rel_l2
from helpers import build_filled_model, build_full_model, build_tiny_model, full_inputs, probe_errors, tiny_inputs
from oracle import cwfa_oracle as O
from oracle.weights import deterministic_fill, seeded_randn

TOL = 2e-5


@pytest.fixture(scope="module")
def small():
    return torch.load(os.path.join(GOLDEN, "r2_small.pt"), weights_only=False)


@pytest.fixture(scope="module")
def ckpt():
    return torch.load(os.path.join(GOLDEN, "r2_ckpt.pt"), weights_only=False)


def build_sequence(fx):
    """The SequenceINN of make_golden_r2.py built from THIS repo's classes (same append calls)."""
    import cwfa_b200.framework as Ff
    import cwfa_b200.modules as Fm
    from cwfa_b200 import networks
    networks.networks_n_chans = 64
    ch, H, W = 6, 12, 16
    torch.manual_seed(0)
    np.random.seed(0)
    seq = Ff.SequenceINN(ch, H, W)
    seq.append(Fm.PermuteRandom, seed=11)
    seq.append(Fm.GLOWCouplingBlock, cond=0, cond_shape=(ch, H, W), subnet_constructor=networks.wavelet_flow_subnetwork2D)
    seq.append(Fm.HaarTransform1D, order_by_wavelet=True)
    seq.append(Fm.ConditionalAffineTransform, cond=1, cond_shape=(ch, H, W), subnet_constructor=networks.wavelet_flow_subnetwork2D)
    seq.append(Fm.HaarDownsampling, order_by_wavelet=True)
    seq.append(Fm.PermuteRandom, seed=12)
    seq.eval()
    sd = deterministic_fill(seq.state_dict(), 600)
    sd.update({k: v.clone() for k, v in seq.state_dict().items() if "haar_weights" in k})
    sd.update({k: v.clone() for k, v in fx["seq/perms"].items()})
    seq.load_state_dict(sd)
    return seq


SEQ_STEPS = [{"type": "perm_chan"}, {"type": "GLOW", "cond": 0}, {"type": "haar1d"}, {"type": "cat", "cond": 1},
             {"type": "haar2d", "order_by_wavelet": True}, {"type": "perm_chan"}]


def test_sequence_inn_structure_and_oracle(small):
    seq = build_sequence(small)
    assert {k: tuple(v.shape) for k, v in seq.state_dict().items()} == small["seq/keys"]          # same keys / shapes as the reference
    assert [tuple(s) for s in seq.shapes] == small["seq/shapes"]
    assert all(torch.equal(seq.state_dict()[k], v) for k, v in small["seq/perms"].items())         # numpy-seeded perms reproduce
    with pytest.raises(ValueError):
        seq.output_dims([(6, 12, 16)])
    sd = {k: v.detach().clone() for k, v in seq.state_dict().items()}
    x = seeded_randn((2, 6, 12, 16), 90)
    cs = [seeded_randn((2, 6, 12, 16), 91), seeded_randn((2, 6, 12, 16), 92)]
    y, j = O.sequence_inn(sd, SEQ_STEPS, x, cs)
    assert rel_l2(y, small["seq/fwd"]) < TOL and rel_l2(j, small["seq/fwd_jac"]) < 1e-5
    xr, jr = O.sequence_inn(sd, SEQ_STEPS, y, cs, rev=True)
    assert rel_l2(xr, x) < 1e-5 and rel_l2(jr, small["seq/rev_jac"]) < 1e-4


def test_lrnn_reads_last_flow_levels_mean_volume(golden_tiny, small):
    """CWFA.py:882: ``cond_nets[L-1](views, mean_vols_cache[n_net-1])`` -- a mean-volume list with one entry per flow level
    (the reference's cache) must route its LAST entry into the LRNN."""
    om = build_tiny_model(golden_tiny).export_for_oracle()
    views, mean_vols = tiny_inputs(golden_tiny)
    L = len(om["levels"])
    outs, jacs = O.reconstruct(om, views, mean_vols[:L], return_all=True)
    assert rel_l2(outs[L], small["refconv/lrnn"]) < TOL
    for n in range(L):
        assert rel_l2(outs[n], small[f"refconv/vol{n}"]) < TOL
        assert abs(float(jacs[n][0] - small[f"refconv/jac{n}"][0])) < 1e-3 * max(1.0, abs(float(jacs[n][0])))
    from cwfa_b200.pipeline import lrnn_mean_volume
    assert lrnn_mean_volume(mean_vols[:L], L) is mean_vols[L - 1]
    assert lrnn_mean_volume(list(mean_vols[:L]) + [None], L) is None


def test_multi_sample_path(golden_tiny, small):
    om = build_tiny_model(golden_tiny).export_for_oracle()
    views, mean_vols = tiny_inputs(golden_tiny)
    n = small["nsamples/level"]
    lv = om["levels"][n]
    K = 3
    vol = small["refconv/lrnn"]
    zs = seeded_randn((K,) + tuple(vol.shape[1:]), small["nsamples/z_seed"], small["nsamples/z_scale"])
    rep = lambda t: t.repeat(K, 1, 1, 1)
    out, jac = O.level_inverse(lv["inn"], lv["spec"], zs, rep(vol), rep(O.cond_network(lv["cond"], views)), rep(mean_vols[n]))
    assert rel_l2(out.mean(0, keepdim=True), small["nsamples/mean"]) < TOL and rel_l2(jac, small["nsamples/jac"]) < 1e-5


def build_dlr_model(golden_tiny, small, device="cpu"):
    cfg = golden_tiny["config"]
    m = build_filled_model(cfg["D"], cfg["S"], cfg["MAX"], small["dlr/specs"], small["dlr/perms"], dict(inn=700, lrnn=300), device,
                           disable_low_res_input=1)
    return m


def test_disable_low_res_input_graph_and_oracle(golden_tiny, small):
    m = build_dlr_model(golden_tiny, small)
    ours = lambda mod: {k: (tuple(v.shape), str(v.dtype)) for k, v in mod.state_dict().items()}
    for n, ref in enumerate(small["dlr/keys"]):
        assert ours(m.conv_inn[n]) == ref
        assert [tuple(d) for d in m.conv_inn[n].dims_c] == small["dlr/dims_c"][n]
    om = m.export_for_oracle()
    assert [lv["spec"] for lv in om["levels"]] == small["dlr/specs"]
    views, mean_vols = tiny_inputs(golden_tiny)
    L = len(om["levels"])
    outs, jacs = O.reconstruct(om, views, list(mean_vols[:L]) + [None], return_all=True, disable_low_res_input=True)
    assert rel_l2(outs[L], small["dlr/lrnn"]) < TOL
    for n in range(L):
        assert rel_l2(outs[n], small[f"dlr/vol{n}"]) < TOL
        assert abs(float(jacs[n][0] - small[f"dlr/jac{n}"][0])) < 1e-3 * max(1.0, abs(float(jacs[n][0])))
    cfg = golden_tiny["config"]
    xg, cg = seeded_randn((2, cfg["D"], cfg["S"], cfg["S"]), 96), seeded_randn((2, cfg["D"] // 2, cfg["S"], cfg["S"]), 97)
    r = O.forward_nll(om, xg, None, None, disable_low_res_input=True, low_res_conditions=[cg])[0]
    assert rel_l2(r["z"], small["dlr/fwd_z"]) < TOL and rel_l2(r["lo"], small["dlr/fwd_lo"]) < TOL
    assert rel_l2(r["logdet"], small["dlr/fwd_jac"]) < 1e-4


SOFTSIGN = lambda u: u / (1.0 + u.abs())
CLAMP_ACTS = {"TANH": "TANH", "SIGMOID": "SIGMOID", "callable": SOFTSIGN}


def build_clamp_block(name, act, device="cpu"):
    import cwfa_b200.modules as Fm
    from cwfa_b200 import networks
    networks.networks_n_chans = 64
    cls = {"cat": Fm.ConditionalAffineTransform, "GLOW": Fm.GLOWCouplingBlock, "RNVP": Fm.RNVPCouplingBlock, "GIN": Fm.GINCouplingBlock}[name]
    torch.manual_seed(0)
    m = cls([(6, 12, 16)], dims_c=[(6, 12, 16)], subnet_constructor=networks.wavelet_flow_subnetwork2D, clamp=1.7, clamp_activation=act).eval()
    m.load_state_dict(deterministic_fill(m.state_dict(), 400))
    return m.to(device)


@pytest.mark.parametrize("tag", list(CLAMP_ACTS))
@pytest.mark.parametrize("name", ["cat", "GLOW", "RNVP", "GIN"])
def test_oracle_clamp_activations(small, tag, name):
    m = build_clamp_block(name, CLAMP_ACTS[tag])
    sd = {"m." + k: v.detach().clone() for k, v in m.state_dict().items()}
    x, c_lf = seeded_randn((2, 6, 12, 16), 23), seeded_randn((2, 6, 12, 16), 24)
    for rev, key in ((False, "fwd"), (True, "rev")):
        if name == "cat":
            y, j = O.cat_block(sd, "m.", x, [c_lf], rev, first=False, clamp=1.7, act=CLAMP_ACTS[tag])
        else:
            y, j = O.glow_block(sd, "m.", x, [c_lf], rev, name, clamp=1.7, act=CLAMP_ACTS[tag])
        assert rel_l2(y, small[f"clamp/{tag}/{name}/{key}"]) < TOL
        assert max_abs(j, small[f"clamp/{tag}/{name}/{key}_jac"]) < 1e-3
    with pytest.raises(ValueError):
        build_clamp_block(name, "RELU")


def build_ckpt_model(ckpt, device="cpu"):
    import cwfa_b200
    from cwfa_b200.data import load_checkpoints
    cfg = ckpt["config"]
    np.random.seed(9)
    torch.manual_seed(9)                                     # deliberately NOT the generator's seed: everything must come from the files
    m = cwfa_b200.CWFAModel(n_depths=cfg["D"], volume_side_size=cfg["S"], INN_max_down_steps=cfg["MAX"], INN_internal_chans=cfg["INN_internal_chans"],
                            INN_cond_chans=cfg["INN_cond_chans"], INN_n_blocks=cfg["INN_n_blocks"])
    stats = load_checkpoints(m, os.path.join(GOLDEN, "ckpt_ref"), permute_dim_axes=ckpt["axes"])
    return m.to(device), stats


def test_reference_written_checkpoint_loads_and_reproduces_reference(ckpt):
    """Files written by the reference's OWN serialize_INN_step (networks.py:708-730) -> load_checkpoints -> the oracle reproduces
    what the reference computed with those networks; training_statistics and the pickled argparse.Namespace come back too."""
    from cwfa_b200.data import load_INN_steps
    path = os.path.join(GOLDEN, "ckpt_ref")
    assert sorted(os.listdir(path)) == ckpt["files"]
    steps = load_INN_steps(path)
    assert sorted(steps) == [1, 2] and all(v[0] == 7 for v in steps.values())
    raw = torch.load(steps[1][1], map_location="cpu", weights_only=False)
    assert raw["args"].INN_down_steps == 1 and raw["args"].INN_internal_chans == ckpt["config"]["INN_internal_chans"]
    m, stats = build_ckpt_model(ckpt)
    assert all(torch.equal(a, b) for a, b in zip(stats, ckpt["training_statistics"]))
    om = m.export_for_oracle()
    cfg = ckpt["config"]
    D, S = cfg["D"], cfg["S"]
    views = seeded_randn((2, 29, S, S), 41)
    for ix in range(cfg["MAX"] - 1):
        lv = om["levels"][ix]
        ch = D // 2 ** (ix + 1)
        lo, mv = seeded_randn((2, ch, S, S), 42 + ix), seeded_randn((2, ch, S, S), 44 + ix, 0.1)
        c0 = O.cond_network(lv["cond"], views)
        assert rel_l2(c0, ckpt[f"step{ix + 1}/cond"]) < TOL
        vol, jac = O.level_inverse(lv["inn"], lv["spec"], torch.zeros_like(lo), lo, c0, mv)
        assert rel_l2(vol, ckpt[f"step{ix + 1}/vol"]) < TOL and rel_l2(jac, ckpt[f"step{ix + 1}/jac"]) < 1e-4
        z, _, jf = O.level_forward(lv["inn"], lv["spec"], vol, c0, mv)
        assert max_abs(z, ckpt[f"step{ix + 1}/z_back"]) < 1e-4 and rel_l2(jf, ckpt[f"step{ix + 1}/jac_fwd"]) < 1e-4


def test_saved_checkpoints_are_readable_by_the_reference_loader_logic(ckpt, tmp_path):
    """save_checkpoints writes an argparse.Namespace the reference's loader can use by ATTRIBUTE (CWFA.py:483-508), from a dict,
    a Namespace or nothing; and the file round-trips."""
    import argparse
    from cwfa_b200.data import load_checkpoints, save_checkpoints
    m, _ = build_ckpt_model(ckpt)
    for k, args in enumerate((None, {"learning_rate": 1e-4}, argparse.Namespace(learning_rate=1e-4, INN_n_blocks=2))):
        d = str(tmp_path / f"c{k}")
        save_checkpoints(m, d, epoch=3, training_statistics=(1.0, 2.0), args=args)
        for step in (1, 2, 3):
            data = torch.load(os.path.join(d, f"model_step_{step}__ep_3"), map_location="cpu", weights_only=False)
            a = data["args"]
            a.INN_down_steps = step                                  # what CWFA.py:489 does to the loaded object
            assert a.INN_internal_chans == 16 and a.INN_use_perm == 1 and a.INN_block_type == "CAT" and a.INN_n_blocks == 2
            assert a.INN_use_bias == 1 and a.INN_max_down_steps == 3 and "force_last_step_NF" not in a
        m2, _ = build_ckpt_model(ckpt)
        with torch.no_grad():
            for p in m2.parameters():
                if p.dtype.is_floating_point:
                    p.zero_()
        assert load_checkpoints(m2, d) == (1.0, 2.0)
        assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
        assert [mod.axis for mod in m2.conv_inn[1].module_list if hasattr(mod, "axis")] == [mod.axis for mod in m.conv_inn[1].module_list if hasattr(mod, "axis")]


# ---- the FULL config: the oracle against the probe of the unmodified reference at 96 x 512 x 512 ------------------------
@pytest.fixture(scope="module")
def full_fx():
    return torch.load(os.path.join(GOLDEN, "r2_full.pt"), weights_only=False)


def test_oracle_full_size_inverse_vs_reference_probe(full_fx):
    """BASELINE.json configs[1] at full size: per-level norms, 4096 seeded sample values and log-dets of the reference."""
    om = build_full_model(full_fx).export_for_oracle()
    views, mean_vols = full_inputs(full_fx)
    with torch.no_grad():
        outs, jacs = O.reconstruct(om, views, mean_vols, bn_mode="batch", return_all=True)
    L = len(om["levels"])
    e = probe_errors(outs[L], full_fx["inv/lrnn"])
    assert e[0] < 1e-5 and e[1] < 1e-4, e
    for n in range(L):
        e = probe_errors(outs[n], full_fx[f"inv/vol{n}"])
        assert e[0] < 1e-5 and e[1] < 1e-4, (n, e)
        r = float(full_fx[f"inv/jac{n}"][0])
        assert abs(float(jacs[n][0]) - r) < 1e-5 * abs(r) + 1e-2, (n, float(jacs[n][0]), r)


def test_oracle_full_size_forward_vs_reference_probe(full_fx):
    """BASELINE.json configs[2] at full size, two of the probe's eight frames on the CPU (the GPU tests cover all eight)."""
    om = build_full_model(full_fx).export_for_oracle()
    _, mean_vols = full_inputs(full_fx)
    seeds = full_fx["config"]["seeds"]
    for b in (0, 5):
        x = seeded_randn((1, 96, 512, 512), seeds["fwd_x"] + b)
        vB = seeded_randn((1, 29, 512, 512), seeds["fwd_views"] + b)
        with torch.no_grad():
            res = O.forward_nll(om, x, vB, mean_vols)
        for n, r in enumerate(res):
            ez, el = probe_errors(r["z"], full_fx[f"fwd/{b}/z{n}"]), probe_errors(r["lo"], full_fx[f"fwd/{b}/lo{n}"])
            assert ez[0] < 1e-5 and ez[1] < 1e-4 and el[1] < 1e-5, (b, n, ez, el)
            rj = float(full_fx[f"fwd/{b}/jac{n}"][0])
            assert abs(float(r["logdet"][0]) - rj) < 1e-5 * abs(rj) + 1e-2
            assert abs(float(r["sumsq"][0]) - float(full_fx[f"fwd/{b}/sumsq{n}"])) < 1e-4 * float(full_fx[f"fwd/{b}/sumsq{n}"])

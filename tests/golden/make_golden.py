"""Generate golden fixtures by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the BUILD container only (the reference does not exist on the GPU box):
    python tests/golden/make_golden.py
Writes tests/golden/tiny.pt and tests/golden/modules.pt (a few hundred KB each).

Weights: the reference model is built with the reference's own constructors
(CWFA.py:478-529 recipe), then every float parameter/buffer is overwritten with
oracle/weights.py:deterministic_fill so the same weights can be regenerated on the GPU box.
Permutations (numpy-seeded, networks.py:343-357) and the PermuteDim axis (not serialised,
INN_utils.py:58-61) are recorded in the fixture.
Pins: z = 0; cond nets .eval(); LRNN U-Net drop_out = 0; BatchNorm mode
stated per case ('batch' = the reference's .train() behaviour, 'running' = .eval()).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference            # noqa: E402
from oracle.weights import deterministic_fill, seeded_randn  # noqa: E402

networks, CWFA, Ff, Fm, INN_utils = import_reference()
torch.set_grad_enabled(False)


def spec_of(inn):
    """Node sequence of the flow branch of one level (what state_dict does not carry)."""
    nodes = []
    mods = list(inn.module_list)
    for i, m in enumerate(mods):
        name = type(m).__name__
        if name in ("HaarTransform1D", "Split"):
            continue
        if name == "ConditionalAffineTransform":
            first = not m.subnet.normal
            nodes.append({"idx": i, "type": "cat_first" if first else "cat"})
        elif name == "PermuteRandom":
            nodes.append({"idx": i, "type": "perm_chan"})
        elif name == "PermuteDim":
            nodes.append({"idx": i, "type": "perm_dim", "axis": int(m.dims_to_permute[1])})
        elif name in ("GLOWCouplingBlock", "GINCouplingBlock", "RNVPCouplingBlock"):
            nodes.append({"idx": i, "type": name[:4].rstrip("C").upper() if name[:3] != "GIN" else "GIN"})
        else:
            raise RuntimeError(name)
    return {"nodes": nodes}


def build_reference(D, S, MAX, seed, block_type="CAT", lrnn_meanvol=True):
    torch.manual_seed(seed)
    np.random.seed(seed)
    inns, conds = [], []
    for ix in range(MAX - 1):
        ctor = lambda ix=ix: networks.cond_network(29, D // 2 ** (ix + 1), ix + 1, MAX, [], 32)
        cn, graphs = networks.conditional_wavelet_flow(
            input_volume_shape=[D, S, S], condition_shape=[1, 29, S, S],
            st_subnet=networks.wavelet_flow_subnetwork2D, conditional_network=ctor,
            n_internal_ch=64, n_down_steps=ix + 1, use_permutations=True,
            block_type=block_type, n_blocks=4, disable_low_res_input=False)
        inns.append(graphs[ix].eval())
        conds.append(cn.eval())
    nd = D // 2 ** (MAX - 1)
    enc = networks.Encoder(29, nd, MAX, 64, 1)
    if S != 512 and lrnn_meanvol:
        enc.net.conv3d = torch.nn.Sequential(networks.ConvNeXt(nd, 64, 0.05, size=S),
                                             networks.ConvNeXt(64, nd, 0.05, size=S))
    enc.net.deconv[1].drop_out = 0.0          # F.dropout2d is otherwise always on (unet.py:80,86)
    for cnx in enc.net.conv3d:
        cnx.drop_prob = 0.0                   # drop_path is active in .train() mode (networks.py:502)
    return inns, conds, enc


def fill(module, seed):
    module.load_state_dict(deterministic_fill(module.state_dict(), seed))


def tiny_case():
    D, S, MAX = 16, 64, 3
    inns, conds, enc = build_reference(D, S, MAX, seed=0)
    for n, (i, c) in enumerate(zip(inns, conds)):
        fill(i, 100 + n)
        fill(c, 200 + n)
    fill(enc, 300)
    views = seeded_randn((1, 29, S, S), 1)
    mean_vols = [seeded_randn((1, D // 2 ** (n + 1), S, S), 10 + n, 0.1) for n in range(MAX - 1)]
    mean_vols.append(seeded_randn((1, D // 2 ** (MAX - 1), S, S), 10 + MAX - 1, 0.1))
    fx = {"config": dict(D=D, S=S, MAX=MAX, seeds=dict(inn=100, cond=200, lrnn=300, views=1, mean=10)),
          "specs": [spec_of(i) for i in inns],
          "perms": [{k: v.clone() for k, v in i.state_dict().items() if "perm" in k} for i in inns]}
    shapes = lambda m: {k: (tuple(v.shape), str(v.dtype)) for k, v in m.state_dict().items()}
    fx["keys"] = {"inn": [shapes(i) for i in inns], "cond": [shapes(c) for c in conds], "lrnn": shapes(enc)}
    # --- inverse reconstruction, CWFA.py:865-924 ---
    for bn_mode in ("batch", "running"):
        enc.train() if bn_mode == "batch" else enc.eval()
        for use_mv in (False, True):
            fill(enc, 300)      # .train() forwards update BN running stats; start every case from the same weights
            vol = enc(views, mean_vols[MAX - 1])[-1] if use_mv else enc(views)[-1]
            tag = f"recon/{bn_mode}/{'mv' if use_mv else 'nomv'}"
            fx[f"{tag}/lrnn"] = vol.clone()
            for n in range(MAX - 2, -1, -1):
                c0 = conds[n](views)[-1].float()
                z = torch.zeros((1,) + tuple(inns[n].global_out_shapes[0]))
                vol, jac = inns[n]([z, vol], c=[c0, mean_vols[n]], rev=True)
                fx[f"{tag}/vol{n}"] = vol.clone()
                fx[f"{tag}/jac{n}"] = jac.clone()
                if bn_mode == "batch" and not use_mv:
                    fx[f"cond{n}"] = c0.clone()
    # --- forward pyramid with real conditions, CWFA.py:966-978 ---
    B = 2
    x = seeded_randn((B, D, S, S), 2)
    vB = seeded_randn((B, 29, S, S), 3)
    for n in range(MAX - 1):
        c0 = conds[n](vB)[-1].float()
        (z, lo), jac = inns[n](x, c=[c0, mean_vols[n].repeat(B, 1, 1, 1)])
        fx[f"fwd/z{n}"], fx[f"fwd/lo{n}"], fx[f"fwd/jac{n}"] = z.clone(), lo.clone(), jac.clone()
        # reference loss formula, CWFA.py:183-189
        fx[f"fwd/nll_ref{n}"] = ((0.5 * torch.norm(z) ** 2 - jac) / lo.numel()).clone()
        # round trip through the reference itself
        xr, jr = inns[n]([z, lo], c=[c0, mean_vols[n].repeat(B, 1, 1, 1)], rev=True)
        fx[f"fwd/roundtrip_err{n}"] = (xr - x).abs().max()
        x = lo
    # --- the reference's own GT-pyramid helper (zero conditions), CWFA.py:134-196 ---
    import types
    ag = types.SimpleNamespace(force_all_steps_NF=0, INN_max_down_steps=MAX)
    gt = seeded_randn((B, D, S, S), 4)
    losses, gt_cache, prior_errors, ljs = CWFA.evaluate_INN_forward(inns, conds + [enc], ag, None, gt.clone(), vB,
                                                                    (0.0, 1.0, 0.0, 1.0, 0.0, 1.0))
    fx["evalfwd/losses"] = torch.stack([l.detach() for l in losses])
    fx["evalfwd/prior"] = torch.stack([l.detach() for l in prior_errors])
    fx["evalfwd/logjac"] = torch.stack([l.detach() for l in ljs])
    fx["evalfwd/gt_last"] = gt_cache[MAX - 1].clone()
    torch.save(fx, os.path.join(HERE, "tiny.pt"))
    print("tiny.pt", {k: (tuple(v.shape) if torch.is_tensor(v) else "...") for k, v in fx.items()})


def module_cases():
    fx = {}
    # HaarTransform1D (INN_utils.py:126-174)
    x = seeded_randn((2, 12, 10, 14), 20)
    m = INN_utils.HaarTransform1D([(12, 10, 14)], order_by_wavelet=True)
    (y,), j = m((x,), rev=False)
    (xr,), jr = m((y,), rev=True)
    fx["haar1d/fwd"], fx["haar1d/rev_of_x"] = y.clone(), m((x,), rev=True)[0][0].clone()
    fx["haar1d/jac"] = torch.tensor([j, jr])
    # FrEIA 2-D Haar (reshapes.py:191-374)
    x = seeded_randn((2, 3, 8, 12), 21)
    for obw in (False, True):
        for reb in (1.0, 0.5):
            m = Fm.HaarDownsampling([(3, 8, 12)], order_by_wavelet=obw, rebalance=reb)
            (y,), j = m((x.clone(),), rev=False)
            fx[f"haar2d/down/{int(obw)}/{reb}"] = y.clone()
            fx[f"haar2d/down_jac/{int(obw)}/{reb}"] = torch.tensor(float(j))
            (xr,), jr = m((y.clone(),), rev=True)
            fx[f"haar2d/up/{int(obw)}/{reb}"] = xr.clone()
            fx[f"haar2d/up_jac/{int(obw)}/{reb}"] = torch.tensor(float(jr))
    # permutations (fixed_transforms.py:11-46, INN_utils.py:46-87)
    x = seeded_randn((2, 6, 8, 8), 22)
    m = Fm.PermuteRandom([(6, 8, 8)], seed=3)
    fx["perm_chan/perm"] = m.perm.data.clone()
    fx["perm_chan/fwd"] = m((x,))[0][0].clone()
    fx["perm_chan/rev"] = m((x,), rev=True)[0][0].clone()
    for trial in range(4):          # the axis is drawn un-seeded; record it
        m = INN_utils.PermuteDim([(6, 8, 8)], seed=5 + trial)
        ax = int(m.dims_to_permute[1])
        fx[f"perm_dim/{trial}/axis"] = torch.tensor(ax)
        fx[f"perm_dim/{trial}/perm"] = m.perm.data.clone()
        fx[f"perm_dim/{trial}/fwd"] = m((x,))[0][0].clone()
        fx[f"perm_dim/{trial}/rev"] = m((x,), rev=True)[0][0].clone()
    # single coupling blocks with the CWFA subnets (coupling_layers.py)
    networks.networks_n_chans = 64
    ch, H, W = 6, 12, 16
    x = seeded_randn((2, ch, H, W), 23)
    c_lf = seeded_randn((2, ch, H, W), 24)
    c_mv = seeded_randn((2, ch, H, W), 25, 0.1)
    cases = {
        "cat": (Fm.ConditionalAffineTransform, networks.wavelet_flow_subnetwork2D, [c_lf]),
        "cat_first": (Fm.ConditionalAffineTransform, networks.wavelet_flow_subnetwork2D_first, [c_mv, c_lf]),
        "GLOW": (Fm.GLOWCouplingBlock, networks.wavelet_flow_subnetwork2D, [c_lf]),
        "GIN": (Fm.GINCouplingBlock, networks.wavelet_flow_subnetwork2D, [c_lf]),
        "RNVP": (Fm.RNVPCouplingBlock, networks.wavelet_flow_subnetwork2D, [c_lf]),
    }
    for name, (cls, sub, conds) in cases.items():
        torch.manual_seed(0)
        m = cls([(ch, H, W)], dims_c=[(ch, H, W)] * len(conds), subnet_constructor=sub).eval()
        m.load_state_dict(deterministic_fill(m.state_dict(), 400))
        (y,), j = m((x,), c=conds, rev=False)
        (xr,), jr = m((x,), c=conds, rev=True)
        j = j if torch.is_tensor(j) else torch.zeros(2) + j
        jr = jr if torch.is_tensor(jr) else torch.zeros(2) + jr
        fx[f"block/{name}/fwd"], fx[f"block/{name}/fwd_jac"] = y.clone(), j.clone()
        fx[f"block/{name}/rev"], fx[f"block/{name}/rev_jac"] = xr.clone(), jr.clone()
    # lenslet view extraction incl. windows clipped by every image border (XLFMDataset.py:212-242)
    import XLFMDataset
    img = seeded_randn((2, 1, 90, 100), 30)
    coords = [[45, 50], [10, 12], [85, 95], [5, 90], [80, 8], [32, 32], [0, 0], [89, 99]]
    fx["extract_views/coords"] = torch.tensor(coords)
    fx["extract_views/out"] = XLFMDataset.XLFMDatasetFull.extract_views(img, coords, [64, 64]).clone()
    # numerical log-det cross-check on a tiny graph (graph_inn.py:369-407)
    torch.manual_seed(0); np.random.seed(0)
    ctor = lambda: networks.cond_network(29, 2, 1, 3, [], 32)
    cn, graphs = networks.conditional_wavelet_flow(
        input_volume_shape=[4, 4, 4], condition_shape=[1, 29, 4, 4],
        st_subnet=networks.wavelet_flow_subnetwork2D, conditional_network=ctor,
        n_internal_ch=64, n_down_steps=1, use_permutations=True, block_type="CAT",
        n_blocks=4, disable_low_res_input=False)
    inn = graphs[0].eval()
    inn.load_state_dict(deterministic_fill(inn.state_dict(), 500))
    xs = seeded_randn((1, 4, 4, 4), 26)
    cs = [seeded_randn((1, 2, 4, 4), 27), seeded_randn((1, 2, 4, 4), 28, 0.1)]
    (z, lo), jac = inn(xs, c=cs)
    fx["numjac/spec"] = spec_of(inn)
    fx["numjac/perms"] = {k: v.clone() for k, v in inn.state_dict().items() if "perm" in k}
    fx["numjac/z"], fx["numjac/lo"], fx["numjac/jac"] = z.clone(), lo.clone(), jac.clone()
    inn.double()
    fx["numjac/jac_numerical"] = inn.log_jacobian_numerical(xs.double(), c=[c.double() for c in cs]).float()
    torch.save(fx, os.path.join(HERE, "modules.pt"))
    print("modules.pt", len(fx), "entries; numjac", fx["numjac/jac"], fx["numjac/jac_numerical"])


if __name__ == "__main__":
    module_cases()
    tiny_case()

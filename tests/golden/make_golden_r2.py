"""Round-2 golden fixtures, produced by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the BUILD container only (the reference does not exist on the GPU box):
    python tests/golden/make_golden_r2.py [full] [small]
Writes
  tests/golden/r2_small.pt     SequenceINN chain; the LRNN fed with mean_vols_cache[L-2] (CWFA.py:882); the
                               disable_low_res_input=1 graph + driver loop (networks.py:331-338, CWFA.py:899-901);
                               TANH / SIGMOID / callable clamp activations (coupling_layers.py:50-60); the multi-sample
                               repeat + mean path (CWFA.py:903-914)
  tests/golden/ckpt_ref/       two step files written by the reference's own serialize_INN_step (networks.py:708-730)
  tests/golden/r2_ckpt.pt      what the reference computes with the networks stored in those files
  tests/golden/r2_full.pt      PROBE of the reference at the FULL config (96 x 512 x 512, 5 steps, BASELINE.json configs[1]
                               and configs[2]): inverse reconstruction at batch 1 and the forward pyramid for 8 frames --
                               per-level norms, seeded sample positions + values, log-dets, sum z^2 (a few hundred KB)

Weights of the probe: reference constructors, then oracle/weights.py:deterministic_fill (regenerable on the GPU box);
permutations / PermuteDim axes are recorded.  Pins as in make_golden.py (z = 0, dropout / drop_path off, BN mode 'batch').
"""
import argparse
import os
import shutil
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference            # noqa: E402
from oracle.weights import deterministic_fill, seeded_randn  # noqa: E402

networks, CWFA, Ff, Fm, INN_utils = import_reference()
torch.set_grad_enabled(False)

sys.path.insert(0, HERE)
from make_golden import build_reference, fill, spec_of       # noqa: E402  (same recipe as round 1)


def probe_positions(numel: int, n: int, seed: int) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randint(0, numel, (n,), generator=g, dtype=torch.int64)


def probe(t: torch.Tensor, n: int, seed: int) -> dict:
    """Norm, mean and n seeded sample values of a tensor (positions are regenerated from the seed by the tests)."""
    flat = t.reshape(-1)
    return {"norm": float(flat.double().norm()), "mean": float(flat.double().mean()), "n": n, "seed": seed,
            "shape": tuple(t.shape), "values": flat[probe_positions(flat.numel(), n, seed)].clone()}


# ---------------------------------------------------------------------------------------------------------------------
def full_probe():
    D, S, MAX = 96, 512, 5
    t0 = time.time()
    inns, conds, enc = build_reference(D, S, MAX, seed=0)
    for n, (i, c) in enumerate(zip(inns, conds)):
        fill(i, 100 + n)
        fill(c, 200 + n)
    fill(enc, 300)
    enc.train()                                              # the reference's LRNN mode (CWFA.py:531-532): batch-statistics BN
    fx = {"config": dict(D=D, S=S, MAX=MAX, seeds=dict(inn=100, cond=200, lrnn=300, views=1, mean=10, fwd_x=1000, fwd_views=2000),
                         n_fwd_frames=8, bn_mode="batch"),
          "specs": [spec_of(i) for i in inns],
          "perms": [{k: v.clone() for k, v in i.state_dict().items() if "perm" in k} for i in inns]}
    views = seeded_randn((1, 29, S, S), 1)
    mean_vols = [seeded_randn((1, D // 2 ** (n + 1), S, S), 10 + n, 0.1) for n in range(MAX - 1)]
    # ---- inverse, exactly the loop of CWFA.py:865-924 at z = 0: the LRNN reads mean_vols_cache[n_net-1] = mean_vols[MAX-2]
    vol = enc(views, mean_vols[MAX - 2])[-1]
    fx["inv/lrnn"] = probe(vol, 4096, 7000)
    for n in range(MAX - 2, -1, -1):
        c0 = conds[n](views)[-1].float()
        z = torch.zeros((1,) + tuple(inns[n].global_out_shapes[0]))
        vol, jac = inns[n]([z, vol], c=[c0, mean_vols[n]], rev=True)
        fx[f"inv/vol{n}"] = probe(vol, 4096, 7001 + n)
        fx[f"inv/jac{n}"] = jac.clone()
        print(f"  inverse level {n}: |vol| {fx[f'inv/vol{n}']['norm']:.4f} jac {float(jac[0]):.3f}  ({time.time() - t0:.0f} s)", flush=True)
    # ---- forward pyramid with real conditions (CWFA.py:966-978), 8 frames run one by one (frames are independent)
    for b in range(8):
        x = seeded_randn((1, D, S, S), 1000 + b)
        vB = seeded_randn((1, 29, S, S), 2000 + b)
        for n in range(MAX - 1):
            c0 = conds[n](vB)[-1].float()
            (z, lo), jac = inns[n](x, c=[c0, mean_vols[n]])
            fx[f"fwd/{b}/z{n}"] = probe(z, 1024, 8000 + 10 * b + n)
            fx[f"fwd/{b}/lo{n}"] = probe(lo, 256, 9000 + 10 * b + n)
            fx[f"fwd/{b}/jac{n}"] = jac.clone()
            fx[f"fwd/{b}/sumsq{n}"] = (z.double() ** 2).sum().float()
            x = lo
        print(f"  forward frame {b} done ({time.time() - t0:.0f} s)", flush=True)
    torch.save(fx, os.path.join(HERE, "r2_full.pt"))
    print("r2_full.pt", os.path.getsize(os.path.join(HERE, "r2_full.pt")), "bytes")


# ---------------------------------------------------------------------------------------------------------------------
def checkpoint_case():
    """Two flow steps of a small config, written to disk by the reference's serialize_INN_step with the argparse.Namespace
    its loader expects (CWFA.py:483-508), plus what the reference computes with those networks."""
    D, S, MAX, NCH, CC = 8, 16, 3, 16, 8
    torch.manual_seed(3)
    np.random.seed(3)
    out_dir = os.path.join(HERE, "ckpt_ref")
    shutil.rmtree(out_dir, ignore_errors=True)
    os.makedirs(out_dir)
    fx = {"config": dict(D=D, S=S, MAX=MAX, INN_internal_chans=NCH, INN_cond_chans=CC, INN_n_blocks=2), "axes": {}}
    stats = (torch.tensor(0.25), torch.tensor(1.5), torch.tensor(0.1), torch.tensor(2.0), torch.tensor(-0.3), torch.tensor(0.7))
    views = seeded_randn((2, 29, S, S), 41)
    nets = []
    for ix in range(MAX - 1):
        ctor = lambda ix=ix: networks.cond_network(29, D // 2 ** (ix + 1), ix + 1, MAX, [], CC)
        cn, graphs = networks.conditional_wavelet_flow(
            input_volume_shape=[D, S, S], condition_shape=[1, 29, S, S], st_subnet=networks.wavelet_flow_subnetwork2D,
            conditional_network=ctor, n_internal_ch=NCH, n_down_steps=ix + 1, use_permutations=True, block_type="CAT",
            n_blocks=2, disable_low_res_input=False)
        nets.append((graphs[ix].eval(), cn.eval()))
    # "trained" weights: perturb the default initialisation ONCE per parameter object so nothing is at a special value (the
    # conditioning nets of all steps share ONE PReLU instance, networks.py:209: after loading every file the reference, too,
    # ends up with the value stored in the last file)
    g = torch.Generator().manual_seed(50)
    seen = set()
    for inn, cn in nets:
        for p in list(inn.parameters()) + list(cn.parameters()):
            if p.dtype.is_floating_point and id(p) not in seen:
                seen.add(id(p))
                p.data += 0.05 * torch.randn(p.shape, generator=g)
    for ix, (inn, cn) in enumerate(nets):
        args = argparse.Namespace(INN_down_steps=ix + 1, INN_internal_chans=NCH, INN_use_perm=1, INN_block_type="CAT", INN_n_blocks=2,
                                  INN_use_bias=1, INN_max_down_steps=MAX, INN_cond_chans=CC, n_depths=D, volume_side_size=S,
                                  force_last_step_NF=0)
        networks.serialize_INN_step(inn, cn, None, stats, args, 7, out_dir)
        ch = D // 2 ** (ix + 1)
        lo = seeded_randn((2, ch, S, S), 42 + ix)
        mv = seeded_randn((2, ch, S, S), 44 + ix, 0.1)
        c0 = cn(views)[-1].float()
        z = torch.zeros((2,) + tuple(inn.global_out_shapes[0]))
        vol, jac = inn([z, lo], c=[c0, mv], rev=True)
        (zf, lof), jf = inn(vol, c=[c0, mv])
        fx[f"step{ix + 1}/cond"], fx[f"step{ix + 1}/vol"], fx[f"step{ix + 1}/jac"] = c0.clone(), vol.clone(), jac.clone()
        fx[f"step{ix + 1}/z_back"], fx[f"step{ix + 1}/jac_fwd"] = zf.clone(), jf.clone()
        fx["axes"][ix + 1] = {i: int(m.dims_to_permute[1]) for i, m in enumerate(inn.module_list) if type(m).__name__ == "PermuteDim"}
    fx["files"] = sorted(os.listdir(out_dir))
    fx["training_statistics"] = stats
    torch.save(fx, os.path.join(HERE, "r2_ckpt.pt"))
    print("ckpt_ref:", {f: os.path.getsize(os.path.join(out_dir, f)) for f in fx["files"]})


# ---------------------------------------------------------------------------------------------------------------------
def small_cases():
    fx = {}
    networks.networks_n_chans = 64
    # ---- SequenceINN (FrEIA/framework/sequence_inn.py:10-99): a conditional chain of the blocks this repo implements
    torch.manual_seed(0)
    np.random.seed(0)
    ch, H, W = 6, 12, 16
    seq = Ff.SequenceINN(ch, H, W)
    seq.append(Fm.PermuteRandom, seed=11)
    seq.append(Fm.GLOWCouplingBlock, cond=0, cond_shape=(ch, H, W), subnet_constructor=networks.wavelet_flow_subnetwork2D)
    seq.append(INN_utils.HaarTransform1D, order_by_wavelet=True)
    seq.append(Fm.ConditionalAffineTransform, cond=1, cond_shape=(ch, H, W), subnet_constructor=networks.wavelet_flow_subnetwork2D)
    seq.append(Fm.HaarDownsampling, order_by_wavelet=True)
    seq.append(Fm.PermuteRandom, seed=12)
    seq.eval()
    sd = deterministic_fill(seq.state_dict(), 600)
    sd.update({k: v.clone() for k, v in seq.state_dict().items() if "perm" in k or "haar_weights" in k})
    seq.load_state_dict(sd)
    x = seeded_randn((2, ch, H, W), 90)
    cs = [seeded_randn((2, ch, H, W), 91), seeded_randn((2, ch, H, W), 92)]
    y, j = seq(x, c=cs)
    xr, jr = seq(y, c=cs, rev=True)
    fx["seq/perms"] = {k: v.clone() for k, v in seq.state_dict().items() if "perm" in k}
    fx["seq/keys"] = {k: tuple(v.shape) for k, v in seq.state_dict().items()}
    fx["seq/fwd"], fx["seq/fwd_jac"], fx["seq/shapes"] = y.clone(), j.clone(), [tuple(s) for s in seq.shapes]
    fx["seq/roundtrip_err"], fx["seq/rev_jac"] = (xr - x).abs().max(), jr.clone()
    seq_t = Ff.SequenceINN(ch, H, W, force_tuple_output=True)
    seq_t.append(Fm.PermuteRandom, seed=11)
    yt, jt = seq_t((x,))
    fx["seq/tuple_out"], fx["seq/tuple_jac"] = yt[0].clone(), float(jt)

    # ---- tiny config: LRNN fed with mean_vols_cache[L-2] (the reference's own convention, CWFA.py:882)
    D, S, MAX = 16, 64, 3
    inns, conds, enc = build_reference(D, S, MAX, seed=0)
    for n, (i, c) in enumerate(zip(inns, conds)):
        fill(i, 100 + n)
        fill(c, 200 + n)
    fill(enc, 300)
    enc.train()
    views = seeded_randn((1, 29, S, S), 1)
    mean_vols = [seeded_randn((1, D // 2 ** (n + 1), S, S), 10 + n, 0.1) for n in range(MAX - 1)]
    vol = enc(views, mean_vols[MAX - 2])[-1]
    fx["refconv/lrnn"] = vol.clone()
    for n in range(MAX - 2, -1, -1):
        c0 = conds[n](views)[-1].float()
        z = torch.zeros((1,) + tuple(inns[n].global_out_shapes[0]))
        vol, jac = inns[n]([z, vol], c=[c0, mean_vols[n]], rev=True)
        fx[f"refconv/vol{n}"], fx[f"refconv/jac{n}"] = vol.clone(), jac.clone()
    # ---- multi-sample path (CWFA.py:903-914): n_samples copies of (z, vol, conditions), mean over the samples at batch 1
    n_samples = 3
    vol = fx["refconv/lrnn"]
    n = MAX - 2
    c0 = conds[n](views)[-1].float()
    zs = seeded_randn((n_samples,) + tuple(inns[n].global_out_shapes[0]), 95, 0.7)
    out, jac = inns[n]([zs, vol.repeat(n_samples, 1, 1, 1)], c=[cc.repeat(n_samples, 1, 1, 1) for cc in (c0, mean_vols[n])], rev=True)
    fx["nsamples/level"], fx["nsamples/z_seed"], fx["nsamples/z_scale"] = n, 95, 0.7
    fx["nsamples/mean"], fx["nsamples/jac"] = out.mean(0).unsqueeze(0).clone(), jac.clone()

    # ---- disable_low_res_input = 1 (networks.py:331-338; driver CWFA.py:899-901: the single condition is the previous volume)
    torch.manual_seed(1)
    np.random.seed(1)
    dinns, dconds = [], []
    for ix in range(MAX - 1):
        ctor = lambda ix=ix: networks.cond_network(29, D // 2 ** (ix + 1), ix + 1, MAX, [], 32)
        cn, graphs = networks.conditional_wavelet_flow(
            input_volume_shape=[D, S, S], condition_shape=[1, 29, S, S], st_subnet=networks.wavelet_flow_subnetwork2D,
            conditional_network=ctor, n_internal_ch=64, n_down_steps=ix + 1, use_permutations=True, block_type="CAT",
            n_blocks=4, disable_low_res_input=True)
        dinns.append(graphs[ix].eval())
        dconds.append(cn.eval())
    for n, i in enumerate(dinns):
        fill(i, 700 + n)
    fx["dlr/specs"] = [spec_of(i) for i in dinns]
    fx["dlr/perms"] = [{k: v.clone() for k, v in i.state_dict().items() if "perm" in k} for i in dinns]
    fx["dlr/keys"] = [{k: (tuple(v.shape), str(v.dtype)) for k, v in i.state_dict().items()} for i in dinns]
    fx["dlr/dims_c"] = [[tuple(d) for d in i.dims_c] for i in dinns]
    fill(enc, 300)
    vol = enc(views)[-1]
    fx["dlr/lrnn"] = vol.clone()
    for n in range(MAX - 2, -1, -1):
        z = torch.zeros((1,) + tuple(dinns[n].global_out_shapes[0]))
        cond = vol                                         # cond_processed = [upsampled_vol]  (CWFA.py:900-901)
        vol, jac = dinns[n]([z, vol], c=[cond], rev=True)
        fx[f"dlr/vol{n}"], fx[f"dlr/jac{n}"] = vol.clone(), jac.clone()
    # forward of the finest level with the same single condition (the training-NLL call, CWFA.py:966)
    xg = seeded_randn((2, D, S, S), 96)
    cg = seeded_randn((2, D // 2, S, S), 97)
    (z, lo), jac = dinns[0](xg, c=[cg])
    fx["dlr/fwd_z"], fx["dlr/fwd_lo"], fx["dlr/fwd_jac"] = z.clone(), lo.clone(), jac.clone()

    # ---- clamp activations (coupling_layers.py:50-60)
    ch, H, W = 6, 12, 16
    x = seeded_randn((2, ch, H, W), 23)
    c_lf = seeded_randn((2, ch, H, W), 24)
    softsign = lambda u: u / (1.0 + u.abs())
    for tag, act in (("TANH", "TANH"), ("SIGMOID", "SIGMOID"), ("callable", softsign)):
        for name, cls in (("cat", Fm.ConditionalAffineTransform), ("GLOW", Fm.GLOWCouplingBlock), ("RNVP", Fm.RNVPCouplingBlock),
                          ("GIN", Fm.GINCouplingBlock)):
            torch.manual_seed(0)
            m = cls([(ch, H, W)], dims_c=[(ch, H, W)], subnet_constructor=networks.wavelet_flow_subnetwork2D, clamp=1.7,
                    clamp_activation=act).eval()
            m.load_state_dict(deterministic_fill(m.state_dict(), 400))
            (y,), j = m((x,), c=[c_lf])
            (xr,), jr = m((x,), c=[c_lf], rev=True)
            j = j if torch.is_tensor(j) else torch.zeros(2) + j
            jr = jr if torch.is_tensor(jr) else torch.zeros(2) + jr
            fx[f"clamp/{tag}/{name}/fwd"], fx[f"clamp/{tag}/{name}/fwd_jac"] = y.clone(), j.clone()
            fx[f"clamp/{tag}/{name}/rev"], fx[f"clamp/{tag}/{name}/rev_jac"] = xr.clone(), jr.clone()
    torch.save(fx, os.path.join(HERE, "r2_small.pt"))
    print("r2_small.pt", len(fx), "entries", os.path.getsize(os.path.join(HERE, "r2_small.pt")), "bytes")


if __name__ == "__main__":
    what = sys.argv[1:] or ["small", "ckpt", "full"]
    if "small" in what:
        small_cases()
    if "ckpt" in what:
        checkpoint_case()
    if "full" in what:
        full_probe()

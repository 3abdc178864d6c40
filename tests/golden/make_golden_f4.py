"""Golden outputs of the reference's ActNorm and AllInOneBlock, produced by the UNMODIFIED FrEIA copy
in /root/reference.  Run in the BUILD container only:  python tests/golden/make_golden_f4.py  -> tests/golden/modules_f4.pt"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference                # noqa: E402
from oracle.weights import deterministic_fill, seeded_randn   # noqa: E402

networks, CWFA, Ff, Fm, INN_utils = import_reference()
torch.set_grad_enabled(False)

AI1_CASES = {
    "hard_softplus_cond": dict(kw=dict(), cond=True),
    "soft_sigmoid_revperm": dict(kw=dict(permute_soft=True, global_affine_type="SIGMOID", reverse_permutation=True, global_affine_init=0.8), cond=True),
    "gin_exp_nocond": dict(kw=dict(gin_block=True, global_affine_type="EXP"), cond=False),
    "householder2": dict(kw=dict(learned_householder_permutation=2, affine_clamping=1.5), cond=True),
}


def main():
    fx = {}
    # ActNorm (invertible_resnet.py:11-85): data-dependent init on the first batch
    x = seeded_randn((3, 6, 8, 10), 80) * 1.7 + 0.4
    m = Fm.ActNorm([(6, 8, 10)])
    (y,), j = m((x,))
    (xr,), jr = m((y,), rev=True)
    fx["actnorm/scale"], fx["actnorm/bias"] = m.scale.data.clone(), m.bias.data.clone()
    fx["actnorm/fwd"], fx["actnorm/jac"], fx["actnorm/rev"], fx["actnorm/rev_jac"] = y.clone(), j.clone(), xr.clone(), jr.clone()
    # AllInOneBlock (all_in_one_block.py:13-271) with the CWFA sub-network
    networks.networks_n_chans = 64
    ch, H, W = 6, 12, 16
    x = seeded_randn((2, ch, H, W), 81)
    c = seeded_randn((2, ch, H, W), 82)
    for name, spec in AI1_CASES.items():
        torch.manual_seed(5); np.random.seed(5)
        conds = [c] if spec["cond"] else []
        m = Fm.AllInOneBlock([(ch, H, W)], dims_c=[(ch, H, W)] * len(conds), subnet_constructor=networks.wavelet_flow_subnetwork2D,
                             **spec["kw"]).eval()
        sd = m.state_dict()
        sd.update(deterministic_fill({k: v for k, v in sd.items() if k.startswith("subnet.")}, 400))
        sd["global_scale"] = sd["global_scale"] + seeded_randn(tuple(sd["global_scale"].shape), 83, 0.3)
        sd["global_offset"] = seeded_randn(tuple(sd["global_offset"].shape), 84, 0.2)
        m.load_state_dict(sd)
        (y,), j = m((x.clone(),), c=conds)
        (xr,), jr = m((x.clone(),), c=conds, rev=True)
        fx[f"ai1/{name}/state"] = {k: v.clone() for k, v in m.state_dict().items() if not k.startswith("subnet.")}
        fx[f"ai1/{name}/fwd"], fx[f"ai1/{name}/fwd_jac"] = y.clone(), j.clone()
        fx[f"ai1/{name}/rev"], fx[f"ai1/{name}/rev_jac"] = xr.clone(), jr.clone()
        print(name, float(y.abs().mean()), j, jr)
    torch.save(fx, os.path.join(HERE, "modules_f4.pt"))
    print("modules_f4.pt", os.path.getsize(os.path.join(HERE, "modules_f4.pt")))


if __name__ == "__main__":
    main()

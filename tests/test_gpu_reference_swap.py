"""GPU: the drop-in claim of INTEGRATION.md section 1, executed.  The UNMODIFIED reference source (staged under oracle/_ref/ by
oracle/make_ref.py) is imported with ``FrEIA.framework`` / ``FrEIA.modules`` resolved to THIS package's modules -- the import
swap a maintainer makes in networks.py:10-15 -- and the reference's OWN ``conditional_wavelet_flow`` (networks.py:264-368),
``wavelet_flow_subnetwork2D(_first)`` and ``cond_network`` build the flow.  The resulting GraphINN (this package's framework
and invertible modules around the reference's torch sub-networks) must reproduce the reference goldens.

Skipped when oracle/_ref is absent (a checkout where /root/reference never existed)."""
import importlib.util
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_l2
from helpers import tiny_inputs
from oracle.weights import deterministic_fill

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REF = os.path.join(ROOT, "oracle", "_ref")


def _import_reference_networks_with_swap():
    """exec the reference's networks.py / INN_utils.py / unet.py in a sandboxed module table where ``FrEIA`` IS cwfa_b200."""
    import cwfa_b200.framework as Ff
    import cwfa_b200.modules as Fm
    saved = {k: sys.modules.get(k) for k in ("FrEIA", "FrEIA.framework", "FrEIA.modules", "INN_utils", "unet", "utils", "networks", "XLFMDataset")}
    freia = types.ModuleType("FrEIA")
    freia.framework, freia.modules = Ff, Fm
    sys.modules.update({"FrEIA": freia, "FrEIA.framework": Ff, "FrEIA.modules": Fm})
    for m in ("matplotlib", "matplotlib.pyplot", "tifffile", "multipagetiff", "lion_pytorch"):
        sys.modules.setdefault(m, MagicMock())
    if "numpy.lib.arraysetops" not in sys.modules:
        shim = types.ModuleType("numpy.lib.arraysetops")
        shim.isin = np.isin
        sys.modules["numpy.lib.arraysetops"] = shim
    mods = {}
    try:
        sys.path.insert(0, REF)
        sys.dont_write_bytecode = True
        for name in ("INN_utils", "unet", "XLFMDataset", "utils", "networks"):
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF, name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            mods[name] = mod
    finally:
        sys.path.remove(REF)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mods["networks"]


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "networks.py")), reason="oracle/_ref not staged (run oracle/make_ref.py)")
def test_reference_flow_builder_runs_on_this_packages_modules(golden_tiny):
    import cwfa_b200.framework as Ff
    import cwfa_b200.modules as Fm
    rn = _import_reference_networks_with_swap()
    assert rn.Ff is Ff and rn.Fm is Fm                              # the reference source now builds with this package's classes
    cfg = golden_tiny["config"]
    D, S, MAX = cfg["D"], cfg["S"], cfg["MAX"]
    torch.manual_seed(0)
    np.random.seed(0)
    inns, conds = [], []
    for ix in range(MAX - 1):
        ctor = lambda ix=ix: rn.cond_network(29, D // 2 ** (ix + 1), ix + 1, MAX, [], 32)
        # the reference's HaarTransform1D / PermuteDim come from ITS INN_utils (star-imported into networks.py); the swap of
        # INTEGRATION.md replaces those two names as well
        rn.HaarTransform1D, rn.PermuteDim = Fm.HaarTransform1D, Fm.PermuteDim
        cn, graphs = rn.conditional_wavelet_flow(
            input_volume_shape=[D, S, S], condition_shape=[1, 29, S, S], st_subnet=rn.wavelet_flow_subnetwork2D,
            conditional_network=ctor, n_internal_ch=64, n_down_steps=ix + 1, use_permutations=True, block_type="CAT", n_blocks=4,
            disable_low_res_input=False)
        inn = graphs[ix]
        assert isinstance(inn, Ff.GraphINN) and any(isinstance(m, Fm.ConditionalAffineTransform) for m in inn.module_list)
        assert type(inn.module_list[2].subnet).__module__ == "networks"          # the reference's own torch sub-network inside our block
        sd = deterministic_fill(inn.state_dict(), cfg["seeds"]["inn"] + ix)
        sd.update({k: v.clone() for k, v in golden_tiny["perms"][ix].items()})
        inn.load_state_dict(sd)
        for node in golden_tiny["specs"][ix]["nodes"]:
            if node["type"] == "perm_dim":
                inn.module_list[node["idx"]].dims_to_permute = [1, node["axis"]]
        cn.load_state_dict(deterministic_fill(cn.state_dict(), cfg["seeds"]["cond"] + ix))
        inns.append(inn.eval().to(DEV))
        conds.append(cn.eval().to(DEV))
    views, mean_vols = tiny_inputs(golden_tiny)
    torch.backends.cudnn.allow_tf32 = False                          # the reference's torch convolutions in true fp32 on the GPU
    torch.backends.cuda.matmul.allow_tf32 = False
    vol = golden_tiny["recon/batch/nomv/lrnn"].to(DEV)              # the LRNN output of the reference; the flow levels are under test here
    with torch.no_grad():
        for n in range(MAX - 2, -1, -1):
            c0 = conds[n](views.to(DEV))[-1].float()                # the reference's cond_network (torch convs on the GPU)
            assert rel_l2(c0, golden_tiny[f"cond{n}"]) < 1e-4
            z = torch.zeros((1,) + tuple(inns[n].global_out_shapes[0]), device=DEV)
            vol, jac = inns[n]([z, vol], c=[c0, mean_vols[n].to(DEV)], rev=True)          # the call of CWFA.py:912
            assert rel_l2(vol, golden_tiny[f"recon/batch/nomv/vol{n}"]) < 1e-4, n
            assert abs(float(jac[0] - golden_tiny[f"recon/batch/nomv/jac{n}"][0])) < 1e-3 * max(1.0, abs(float(jac[0])))

"""Import the UNMODIFIED reference (pvjosue/CWFA at /root/reference) in this container.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py (fixture generation) and by
oracle validation here in the build container.  /root/reference does not exist on the GPU
box, so nothing that runs there imports this file.

The reference needs a few non-arithmetic third-party modules that are not installed
(matplotlib, tifffile, multipagetiff, lion_pytorch) and one numpy-1 path
(numpy.lib.arraysetops).  None of them carries arithmetic of the hot path; they are stubbed.
Recipe: SURVEY.md section 8(c).
"""
import os
import sys
import types
from unittest.mock import MagicMock

_HERE = os.path.dirname(os.path.abspath(__file__))
_STAGED = os.path.join(_HERE, "_ref")          # verbatim copy made by oracle/make_ref.py (git-ignored; travels to the GPU box)


def _pick_root() -> str:
    env = os.environ.get("CWFA_REFERENCE_ROOT")
    for cand in (env, "/root/reference", _STAGED):
        if cand and os.path.isfile(os.path.join(cand, "networks.py")):
            return cand
    return env or "/root/reference"


REF_ROOT = _pick_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "networks.py"))


def build_reference_model(D, S, MAX, seed=0, block_type="CAT", n_blocks=4, disable_low_res_input=False, lrnn_size=None):
    """The reference's own model assembly (CWFA.py:478-529): per step ``conditional_wavelet_flow(..., n_down_steps=ix+1)`` keeping
    graph ``ix`` + its ``cond_network``, and the ``Encoder`` for the last step.  Returns (inns, conds, encoder), all .eval();
    the encoder's always-on dropout2d / drop_path are pinned to 0.  ``lrnn_size``: rebuild the
    mean-volume branch for a side length other than the hard-coded 512 (networks.py:472,490)."""
    import numpy as np
    import torch
    networks = import_reference()[0]
    torch.manual_seed(seed)
    np.random.seed(seed)
    inns, conds = [], []
    for ix in range(MAX - 1):
        ctor = lambda ix=ix: networks.cond_network(29, D // 2 ** (ix + 1), ix + 1, MAX, [], 32)
        cn, graphs = networks.conditional_wavelet_flow(
            input_volume_shape=[D, S, S], condition_shape=[1, 29, S, S],
            st_subnet=networks.wavelet_flow_subnetwork2D, conditional_network=ctor,
            n_internal_ch=64, n_down_steps=ix + 1, use_permutations=True,
            block_type=block_type, n_blocks=n_blocks, disable_low_res_input=disable_low_res_input)
        inns.append(graphs[ix].eval())
        conds.append(cn.eval())
    nd = D // 2 ** (MAX - 1)
    enc = networks.Encoder(29, nd, MAX, 64, 1)
    if lrnn_size is not None and lrnn_size != 512:
        enc.net.conv3d = torch.nn.Sequential(networks.ConvNeXt(nd, 64, 0.05, size=lrnn_size), networks.ConvNeXt(64, nd, 0.05, size=lrnn_size))
    enc.net.deconv[1].drop_out = 0.0
    for cnx in enc.net.conv3d:
        cnx.drop_prob = 0.0
    return inns, conds, enc


def reference_inverse(inns, conds, enc, views, mean_vols):
    """The reference's inverse driver loop at z = 0 (CWFA.py:865-924), verbatim in structure: LRNN with mean_vols_cache[n_net-1],
    then ``conv_inn[n]([z, vol], c=[cond_net(views), mean_vols_cache[n]], rev=True)`` from the coarsest level up."""
    import torch
    L = len(inns)
    vol = enc(views, mean_vols[L - 1])[-1] if mean_vols[L - 1] is not None else enc(views)[-1]
    for n in range(L - 1, -1, -1):
        cond_processed = [conds[n](views)[-1].float(), mean_vols[n]]
        z = torch.zeros((views.shape[0],) + tuple(inns[n].global_out_shapes[0]))
        vol, _ = inns[n]([z, vol], c=cond_processed, rev=True)
    return vol


def import_reference():
    """Returns (networks, CWFA, Ff, Fm, INN_utils) modules of the reference."""
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    sys.dont_write_bytecode = True
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm",
              "tifffile", "multipagetiff", "lion_pytorch"):
        if m not in sys.modules:
            sys.modules[m] = MagicMock()
    if "numpy.lib.arraysetops" not in sys.modules:
        import numpy as np
        shim = types.ModuleType("numpy.lib.arraysetops")
        shim.isin = np.isin
        sys.modules["numpy.lib.arraysetops"] = shim
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import FrEIA.framework as Ff
    import FrEIA.modules as Fm
    import INN_utils
    import networks
    try:
        import CWFA
    except Exception:  # CWFA.py pulls in plotting/tensorboard; not needed for arithmetic
        CWFA = None
    return networks, CWFA, Ff, Fm, INN_utils

"""CPU oracle for the CWFA conditional-wavelet-flow hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this file.  Nothing under
cwfa_b200/ imports it; the product path fails loudly when its CUDA library is missing.

What it is: a functional restatement, on CPU tensors, of the arithmetic of the reference
path (pvjosue/CWFA).  It works on plain ``state_dict``s that use the reference's own key
names plus a small ``spec`` dict describing the node sequence of one flow level (the
things the reference does not serialise: PermuteDim axis, block type).  Every function
cites the reference file:line it restates.  Convolutions etc. use torch's CPU functional
ops -- the same ATen ops the reference itself runs on CPU -- so the oracle follows the
reference bit-for-bit in fp32 wherever the op order is the same.

Parity pinning: the reference ships NO tests, golden vectors or fixtures
(SURVEY.md section 4 / 8c), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF:
tests/golden/make_golden.py imports the unmodified reference in the build container,
runs the tiny config and a set of per-module cases, and commits inputs, state_dicts and
outputs under tests/golden/.  tests/test_oracle_golden.py replays them through this file.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

CLAMP = 2.0          # FrEIA/modules/coupling_layers.py:17 (clamp default)
K_ATAN = 0.636       # FrEIA/modules/coupling_layers.py:52 (NOT 2/pi)
INV_SQRT2 = 1.0 / math.sqrt(2.0)


# ----------------------------------------------------------------------------------------
# Haar transforms
# ----------------------------------------------------------------------------------------
def haar1d(x: Tensor, rev: bool = False) -> Tuple[Tensor, float]:
    """Depth-wise (channel axis) Haar DWT / IDWT.  INN_utils.py:142-161.

    fwd: out[:, i] = (x[:, 2i] + x[:, 2i+1]) / sqrt2 ; out[:, h+i] = (x[:, 2i] - x[:, 2i+1]) / sqrt2
    rev: out[:, 2i] = (x[:, i] + x[:, h+i]) / sqrt2 ; out[:, 2i+1] = (x[:, i] - x[:, h+i]) / sqrt2
    log-det is exactly 0 for rebalance=1 (INN_utils.py:135-140).
    """
    h = x.shape[1] // 2
    out = torch.zeros_like(x)
    if not rev:
        out[:, :h] = x[:, ::2] + x[:, 1::2]
        out[:, h:] = x[:, ::2] - x[:, 1::2]
    else:
        out[:, ::2] = x[:, :h] + x[:, h:]
        out[:, 1::2] = x[:, :h] - x[:, h:]
    return out * INV_SQRT2, 0.0


def _haar2d_weights(c: int, dtype) -> Tensor:
    """FrEIA/modules/reshapes.py:240-252."""
    w = torch.ones(4, 1, 2, 2, dtype=dtype)
    w[1, 0, 0, 1] = -1
    w[1, 0, 1, 1] = -1
    w[2, 0, 1, 0] = -1
    w[2, 0, 1, 1] = -1
    w[3, 0, 1, 0] = -1
    w[3, 0, 0, 1] = -1
    return torch.cat([w] * c, 0)


def haar2d(x: Tensor, rev: bool = False, order_by_wavelet: bool = False,
           rebalance: float = 1.0) -> Tuple[Tensor, float]:
    """FrEIA 2-D HaarDownsampling (rev=False) / its inverse.  reshapes.py:273-300.

    Input is NOT mutated (the reference mutates it in rev when order_by_wavelet=False,
    reshapes.py:297 -- a bug we deliberately do not replicate).
    """
    fac_fwd = 0.5 * rebalance
    fac_rev = 0.5 / rebalance
    ndims = x[0].numel()
    if not rev:
        c = x.shape[1]
        jac = ndims * (math.log(16.0) + 4 * math.log(fac_fwd)) / 4.0
        out = F.conv2d(x, _haar2d_weights(c, x.dtype), None, stride=2, groups=c)
        if order_by_wavelet:
            perm = torch.tensor([i + 4 * j for i in range(4) for j in range(c)])
            out = out[:, perm]
        return out * fac_fwd, jac
    c = x.shape[1] // 4
    jac = ndims * (math.log(16.0) + 4 * math.log(fac_rev)) / 4.0
    if order_by_wavelet:
        perm = torch.tensor([i + 4 * j for i in range(4) for j in range(c)])
        perm_inv = torch.empty_like(perm)
        perm_inv[perm] = torch.arange(4 * c)
        x = x[:, perm_inv]
    out = F.conv_transpose2d(x * fac_rev, _haar2d_weights(c, x.dtype), stride=2, groups=c)
    return out, jac


# ----------------------------------------------------------------------------------------
# Permutations
# ----------------------------------------------------------------------------------------
def permute_random(x: Tensor, perm: Tensor, perm_inv: Tensor, rev: bool = False) -> Tensor:
    """Channel gather.  FrEIA/modules/fixed_transforms.py:37-41."""
    return x[:, perm_inv if rev else perm]


def permute_dim(x: Tensor, perm: Tensor, perm_inv: Tensor, axis: int, rev: bool = False) -> Tensor:
    """Row (axis=2) or column (axis=3) gather.  INN_utils.py:73-81."""
    p = perm_inv if rev else perm
    return x.index_select(axis, p)


# ----------------------------------------------------------------------------------------
# Coupling sub-network (the conv trunk that predicts s, t)
# ----------------------------------------------------------------------------------------
def subnet(sd: SD, pre: str, inp: Tensor, normal: bool) -> Tensor:
    """wavelet_flow_subnetwork.forward, networks.py:641-671 (2-D variants :673-706).

    normal=True  (wavelet_flow_subnetwork2D):       b1 = block12(inp),  out = block72(b6)
    normal=False (wavelet_flow_subnetwork2D_first): low, cond = inp[:, :-n], inp[:, -n:]
                 b1 = block1(cond), out = cat(block7(b6), -low / sqrt2)
    """
    def conv(name, t, pad):
        return F.conv2d(t, sd[pre + name + ".weight"], sd[pre + name + ".bias"], padding=pad)

    low = None
    if normal:
        b1 = conv("block12", inp, 0)
    else:
        n = sd[pre + "block1.weight"].shape[1]          # c_in // 2
        low, cond = inp[:, :-n], inp[:, -n:]
        b1 = conv("block1", cond, 0)
    b = b1
    for blk in ("block2", "block4", "block6"):
        t = conv(blk + ".0", b, 1)
        t = F.elu(t)
        t = conv(blk + ".2", t, 0)
        b = t + b
        if blk != "block6":
            b = F.elu(b)
    if normal:
        return conv("block72.1", F.elu(b), 1)
    b7 = conv("block7.1", F.elu(b), 1)
    return torch.cat((b7, -low / math.sqrt(2)), 1)


def f_clamp(u: Tensor, act="ATAN") -> Tensor:
    """Clamp activations of _BaseCouplingBlock, coupling_layers.py:50-60: ATAN -> 0.636 atan(u), TANH -> tanh(u),
    SIGMOID -> 2 (sigmoid(u) - 0.5); anything else is the user's callable."""
    if act == "ATAN":
        return K_ATAN * torch.atan(u)
    if act == "TANH":
        return torch.tanh(u)
    if act == "SIGMOID":
        return 2.0 * (torch.sigmoid(u) - 0.5)
    return act(u)


def affine(x: Tensor, a: Tensor, rev: bool, clamp: float = CLAMP, act="ATAN") -> Tuple[Tensor, Tensor]:
    """s = clamp * f_clamp(a[:, :ch]) (0.636 atan by default); t = a[:, ch:]; coupling_layers.py:490-500.

    fwd: y = exp(s) * x + t, J = +sum(s);  rev: y = (x - t) * exp(-s), J = -sum(s).
    """
    ch = x.shape[1]
    s, t = a[:, :ch], a[:, ch:]
    s = clamp * f_clamp(s, act)
    j = torch.sum(s, dim=tuple(range(1, x.dim())))
    if rev:
        return (x - t) * torch.exp(-s), -j
    return torch.exp(s) * x + t, j


def cat_block(sd: SD, pre: str, x: Tensor, conds: Sequence[Tensor], rev: bool,
              first: bool, clamp: float = CLAMP, act="ATAN") -> Tuple[Tensor, Tensor]:
    """ConditionalAffineTransform.forward, coupling_layers.py:475-500."""
    cond = torch.cat(list(conds), 1) if len(conds) > 1 else conds[0]
    a = subnet(sd, pre + "subnet.", cond, normal=not first)
    return affine(x, a, rev, clamp, act)


def glow_block(sd: SD, pre: str, x: Tensor, conds: Sequence[Tensor], rev: bool,
               kind: str = "GLOW", clamp: float = CLAMP, act="ATAN") -> Tuple[Tensor, Tensor]:
    """_BaseCouplingBlock.forward + GLOW/GIN/RNVP couplings.  coupling_layers.py:62-87,
    :160-229 (RNVP), :232-302 (GLOW), :305-381 (GIN)."""
    l1 = x.shape[1] // 2
    l2 = x.shape[1] - l1
    x1, x2 = x[:, :l1], x[:, l1:]
    nd = tuple(range(1, x.dim()))

    def st(which, u, n_out):
        if kind == "RNVP":
            s = subnet(sd, f"{pre}subnet_s{which}.", u, True)
            t = subnet(sd, f"{pre}subnet_t{which}.", u, True)
        else:
            a = subnet(sd, f"{pre}subnet{which}.", u, True)
            s, t = a[:, :n_out], a[:, n_out:]
        s = clamp * f_clamp(s, act)
        if kind == "GIN":
            s = s - s.mean(1, keepdim=True)
            return s, t, 0.0
        return s, t, torch.sum(s, dim=nd)

    def c1(x1_, u2, r):     # uses subnet2 (coupling_layers.py:268-289)
        s, t, j = st(2, u2, l1)
        return ((x1_ - t) * torch.exp(-s), -j) if r else (torch.exp(s) * x1_ + t, j)

    def c2(x2_, u1, r):     # uses subnet1 (coupling_layers.py:291-302)
        s, t, j = st(1, u1, l2)
        return ((x2_ - t) * torch.exp(-s), -j) if r else (torch.exp(s) * x2_ + t, j)

    cc = list(conds)
    if not rev:
        y1, j1 = c1(x1, torch.cat([x2, *cc], 1), False)
        y2, j2 = c2(x2, torch.cat([y1, *cc], 1), False)
    else:
        y2, j2 = c2(x2, torch.cat([x1, *cc], 1), True)
        y1, j1 = c1(x1, torch.cat([y2, *cc], 1), True)
    j = j1 + j2
    if not torch.is_tensor(j):
        j = torch.zeros(x.shape[0], dtype=x.dtype) + j
    return torch.cat((y1, y2), 1), j


# ----------------------------------------------------------------------------------------
# One flow level (GraphINN of networks.py:305-366)
# ----------------------------------------------------------------------------------------
def _run_node(sd: SD, node: dict, x: Tensor, c_lf: Tensor, c_mean: Optional[Tensor], rev: bool):
    i = node["idx"]
    pre = f"module_list.{i}."
    kind = node["type"]
    zero = torch.zeros(x.shape[0], dtype=x.dtype)
    if kind == "cat_first":
        # node conditions = [Condition (mean-vol delta), Condition I (LF)], networks.py:329-339
        return cat_block(sd, pre, x, [c_mean, c_lf], rev, first=True)
    if kind == "cat":
        return cat_block(sd, pre, x, [c_lf], rev, first=False)
    if kind in ("GLOW", "GIN", "RNVP"):
        return glow_block(sd, pre, x, [c_lf], rev, kind)
    if kind == "perm_chan":
        return permute_random(x, sd[pre + "perm"], sd[pre + "perm_inv"], rev), zero
    if kind == "perm_dim":
        return permute_dim(x, sd[pre + "perm"], sd[pre + "perm_inv"], node["axis"], rev), zero
    raise ValueError(kind)


def level_forward(sd: SD, spec: dict, x: Tensor, c_lf: Tensor, c_mean: Optional[Tensor]):
    """GraphINN.forward(rev=False) of one level: returns (z, lo, logdet[B]).
    graph_inn.py:242-326; node order networks.py:305-366."""
    y, _ = haar1d(x, rev=False)
    h = y.shape[1] // 2
    lo, hi = y[:, :h], y[:, h:]                           # Split, graph_topology.py:73-80
    jac = torch.zeros(x.shape[0], dtype=x.dtype)
    for node in spec["nodes"]:
        hi, j = _run_node(sd, node, hi, c_lf, c_mean, rev=False)
        jac = jac + j
    return hi, lo, jac


def level_inverse(sd: SD, spec: dict, z: Tensor, lo: Tensor, c_lf: Tensor, c_mean: Optional[Tensor]):
    """GraphINN.forward(rev=True): returns (x, logdet[B])."""
    hi = z
    jac = torch.zeros(z.shape[0], dtype=z.dtype)
    for node in reversed(spec["nodes"]):
        hi, j = _run_node(sd, node, hi, c_lf, c_mean, rev=True)
        jac = jac + j
    x, _ = haar1d(torch.cat((lo, hi), 1), rev=True)
    return x, jac


# ----------------------------------------------------------------------------------------
# Conditioning network (networks.py:165-242)
# ----------------------------------------------------------------------------------------
def cond_network(sd: SD, views: Tensor) -> Tensor:
    """cond_network.forward -> ResidualBlock.forward (eval mode: Dropout3d = identity).
    networks.py:195-196, :229-242.  The PReLU module is one shared instance
    (default-arg nn.PReLU(), networks.py:209) stored under three keys."""
    p = "subnetworks.0."
    out = F.conv2d(views, sd[p + "conv1.0.weight"], sd[p + "conv1.0.bias"], padding=1)
    out = F.prelu(out, sd[p + "conv1.1.weight"])
    out = F.conv2d(out, sd[p + "conv2.0.weight"], sd[p + "conv2.0.bias"], padding=1)
    res = F.conv2d(views, sd[p + "downsample.0.weight"], sd[p + "downsample.0.bias"], padding=1)
    out = F.prelu(out + res, sd[p + "relu.weight"])
    v = out.permute(0, 2, 3, 1).unsqueeze(1)                 # (B,1,H,W,ch)
    v = F.conv3d(v, sd[p + "conv3d.0.weight"], sd[p + "conv3d.0.bias"], padding=1)
    v = F.prelu(v, sd[p + "conv3d.1.weight"])
    v = F.conv3d(v, sd[p + "conv3d.3.weight"], sd[p + "conv3d.3.bias"], padding=1)
    return v[:, 0].permute(0, 3, 1, 2).contiguous()


# ----------------------------------------------------------------------------------------
# LRNN (networks.py:505-584, unet.py)
# ----------------------------------------------------------------------------------------
def _bn(sd: SD, pre: str, x: Tensor, mode: str, eps: float = 1e-5) -> Tensor:
    """nn.BatchNorm2d.  mode='batch' = training-mode batch statistics (what the reference
    runs at inference, CWFA.py:531-532); mode='running' = eval-mode running stats."""
    w, b = sd[pre + "weight"], sd[pre + "bias"]
    if mode == "batch":
        return F.batch_norm(x, None, None, w, b, training=True, eps=eps)
    return F.batch_norm(x, sd[pre + "running_mean"], sd[pre + "running_var"], w, b,
                        training=False, eps=eps)


def _unet_block(sd: SD, pre: str, x: Tensor, bn_mode: str) -> Tensor:
    """UNetConvBlock: [conv3x3, PReLU, BN] x2.  unet.py:94-113."""
    for ci, ai, bi in ((0, 1, 2), (3, 4, 5)):
        x = F.conv2d(x, sd[f"{pre}block.{ci}.weight"], sd.get(f"{pre}block.{ci}.bias"), padding=1)
        x = F.prelu(x, sd[f"{pre}block.{ai}.weight"])
        x = _bn(sd, f"{pre}block.{bi}.", x, bn_mode)
    return x


def unet(sd: SD, pre: str, x: Tensor, bn_mode: str = "batch") -> Tensor:
    """UNet.forward with drop_out pinned to 0 (the reference's F.dropout2d is always
    active with p=0.005, unet.py:80,86 -- stochastic, so parity pins p=0).  unet.py:72-91."""
    depth = 1 + max(int(k[len(pre) + 10:].split(".")[0]) for k in sd if k.startswith(pre + "down_path."))
    blocks = []
    for i in range(depth):
        x = _unet_block(sd, f"{pre}down_path.{i}.", x, bn_mode)
        if i != depth - 1:
            blocks.append(x)
            x = F.adaptive_max_pool2d(x, x.shape[-1] // 2)
    for i in range(depth - 1):
        p = f"{pre}up_path.{i}."
        up = F.conv_transpose2d(x, sd[p + "up.weight"], sd.get(p + "up.bias"), stride=2)
        x = _unet_block(sd, p + "conv_block.", up + blocks[-i - 1], bn_mode)     # skip ADD, unet.py:190
    x = F.conv2d(x, sd[pre + "last.0.weight"], sd.get(pre + "last.0.bias"))
    return F.prelu(x, sd[pre + "last.1.weight"])


def convnext(sd: SD, pre: str, x: Tensor) -> Tensor:
    """ConvNeXt.forward in eval mode (drop_path = identity).  networks.py:486-503."""
    up = F.conv2d(x, sd[pre + "input.weight"], sd[pre + "input.bias"])
    m = F.conv2d(up, sd[pre + "m.0.weight"], sd[pre + "m.0.bias"], padding=3)
    m = F.layer_norm(m, m.shape[1:], sd[pre + "m.1.weight"], sd[pre + "m.1.bias"], 1e-5)
    m = F.conv2d(m, sd[pre + "m.2.weight"], sd[pre + "m.2.bias"])
    m = F.gelu(m)
    return m + up


def global_attention(sd: SD, pre: str, x: Tensor) -> Tensor:
    """GlobalAttention.forward: Conv1d(k=3) over the FLATTENED H*W axis.  networks.py:250-262."""
    f = x.reshape(x.shape[0], x.shape[1], -1)
    f = F.conv1d(f, sd[pre + "m.0.weight"], sd[pre + "m.0.bias"], padding=1)
    f = F.relu(f)
    f = F.conv1d(f, sd[pre + "m.2.weight"], sd[pre + "m.2.bias"])
    return torch.sigmoid(f).reshape(x.shape)


def lrnn(sd: SD, views: Tensor, mean_vol: Optional[Tensor] = None, bn_mode: str = "batch") -> Tensor:
    """Encoder.forward -> LRNN.forward.  networks.py:573-584, :544-555."""
    p = "net."
    x = F.conv2d(views, sd[p + "deconv.0.weight"], sd.get(p + "deconv.0.bias"))
    x = unet(sd, p + "deconv.1.", x, bn_mode)
    if mean_vol is not None:
        mp = convnext(sd, p + "conv3d.1.", convnext(sd, p + "conv3d.0.", mean_vol))
        x = x + mp * 2 * (global_attention(sd, p + "attention_3d.", mean_vol) - 0.5)
    return x


# ----------------------------------------------------------------------------------------
# Whole-pipeline drivers
# ----------------------------------------------------------------------------------------
def reconstruct(model: dict, views: Tensor, mean_vols: Sequence[Optional[Tensor]],
                zs: Optional[Sequence[Tensor]] = None, bn_mode: str = "batch",
                return_all: bool = False, disable_low_res_input: bool = False, n_samples: int = 1):
    """Inverse reconstruction, CWFA.py:865-924: LRNN low-res volume, then each flow level
    n = L-1 .. 0 with z = 0 (INN_z_temperature = 0, CWFA.py:906-907).

    model = {"levels": [{"inn": sd, "cond": sd, "spec": spec}, ...], "lrnn": sd}
    mean_vols[n] is the mean-volume delta condition of level n (n < L).  The LRNN receives
    mean_vols[L-1] -- CWFA.py:882 passes mean_vols_cache[n_net-1], the last flow level's condition --
    unless an explicit extra entry mean_vols[L] (tensor, or None = no mean-volume branch) is given.
    disable_low_res_input: the level's single condition is the previous up-sampled volume (CWFA.py:899-901).
    n_samples > 1 (batch 1): n_samples copies of (z, low-res volume, conditions), mean over the samples (CWFA.py:903-914).
    """
    levels = model["levels"]
    L = len(levels)
    mv_last = mean_vols[L] if len(mean_vols) > L else mean_vols[L - 1]
    vol = lrnn(model["lrnn"], views, mv_last, bn_mode)
    outs = {L: vol}
    jacs = {}
    for n in range(L - 1, -1, -1):
        lv = levels[n]
        if disable_low_res_input:
            c_lf, c_mean = vol, None                          # cond_processed = [upsampled_vol]
        else:
            c_lf, c_mean = cond_network(lv["cond"], views), mean_vols[n]
        if n_samples > 1:
            rep = lambda t: None if t is None else t.repeat(n_samples, 1, 1, 1)
            vol, c_lf, c_mean = rep(vol), rep(c_lf), rep(c_mean)
        z = torch.zeros_like(vol) if zs is None or zs[n] is None else zs[n]
        vol, jac = level_inverse(lv["inn"], lv["spec"], z, vol, c_lf, c_mean)
        if n_samples > 1:
            vol = vol.mean(0).unsqueeze(0)
        outs[n] = vol
        jacs[n] = jac
    return (outs, jacs) if return_all else vol


def forward_nll(model: dict, volume: Tensor, views: Tensor, mean_vols: Sequence[Tensor],
                disable_low_res_input: bool = False, low_res_conditions: Optional[Sequence[Tensor]] = None):
    """Forward pyramid with REAL conditions + per-level NLL (CWFA.py:966-978; pyramid
    structure of evaluate_INN_forward, CWFA.py:156-196).

    Returns a list per level of dicts: z, lo, logdet[B], sumsq[B],
    nll_per_sample[B] = (0.5*sumsq_b - logdet_b) / (ch*P),
    nll_ref = the reference's batch-coupled formula (0.5*||Z||^2 - logdet) / Z[-1].numel()
    (CWFA.py:183-189: ||Z||^2 over the WHOLE batch, numel of the lo tensor incl. batch).
    disable_low_res_input: single condition = low_res_conditions[n], default the volume's own low-resolution half.
    """
    res = []
    x = volume
    for n, lv in enumerate(model["levels"]):
        if disable_low_res_input:
            given = low_res_conditions[n] if low_res_conditions is not None and n < len(low_res_conditions) else None
            c_lf = given if given is not None else haar1d(x)[0][:, :x.shape[1] // 2]
            c_mean = None
        else:
            c_lf, c_mean = cond_network(lv["cond"], views), mean_vols[n]
        z, lo, jac = level_forward(lv["inn"], lv["spec"], x, c_lf, c_mean)
        sumsq = (z.double() ** 2).flatten(1).sum(1).to(z.dtype)
        per = (0.5 * sumsq - jac) / z[0].numel()
        ref = (0.5 * torch.norm(z) ** 2 - jac) / lo.numel()
        res.append(dict(z=z, lo=lo, logdet=jac, sumsq=sumsq, nll_per_sample=per, nll_ref=ref))
        x = lo
    return res


def sequence_inn(sd: SD, steps: Sequence[dict], x: Tensor, conds: Sequence[Tensor], rev: bool = False):
    """SequenceINN.forward, FrEIA/framework/sequence_inn.py:68-99: modules applied in order (reversed for rev), log-dets summed.
    ``steps[i]`` = {"type": perm_chan | GLOW | GIN | RNVP | cat | haar1d | haar2d, "cond": index or None, ...}; the
    parameters of module i live under ``module_list.{i}.`` as in the reference's state_dict."""
    jac = 0
    order = range(len(steps))
    for i in (reversed(order) if rev else order):
        st = steps[i]
        pre = f"module_list.{i}."
        c = [] if st.get("cond") is None else [conds[st["cond"]]]
        t = st["type"]
        if t == "perm_chan":
            x, j = permute_random(x, sd[pre + "perm"], sd[pre + "perm_inv"], rev), 0.0
        elif t in ("GLOW", "GIN", "RNVP"):
            x, j = glow_block(sd, pre, x, c, rev, t, st.get("clamp", CLAMP), st.get("act", "ATAN"))
        elif t == "cat":
            x, j = cat_block(sd, pre, x, c, rev, first=False, clamp=st.get("clamp", CLAMP), act=st.get("act", "ATAN"))
        elif t == "haar1d":
            x, j = haar1d(x, rev)
        elif t == "haar2d":
            x, j = haar2d(x, rev, st.get("order_by_wavelet", False), st.get("rebalance", 1.0))
        else:
            raise ValueError(t)
        jac = j + jac
    return x, jac


# ----------------------------------------------------------------------------------------
# Either side of the path: lenslet crop before it, GT pyramid helper
# ----------------------------------------------------------------------------------------
def extract_views(image: Tensor, lenslet_coords, subimage_shape) -> Tensor:
    """XLFMDatasetFull.extract_views, XLFMDataset.py:212-242: crop an S0 x S1 window around every lenslet centre,
    clipped to the image, written bottom/right aligned into the view."""
    h0, h1 = subimage_shape[0] // 2, subimage_shape[1] // 2
    out = torch.zeros((image.shape[0], len(lenslet_coords), subimage_shape[0], subimage_shape[1]), dtype=image.dtype)
    for n, c in enumerate(lenslet_coords):
        cy, cx = int(c[0]), int(c[1])
        ly, lx = max(cy - h0, 0), max(cx - h1, 0)
        patch = image[:, 0, ly:cy + h0, lx:cx + h1]
        out[:, n, subimage_shape[0] - patch.shape[1]:, subimage_shape[1] - patch.shape[2]:] = patch
    return out


def evaluate_inn_forward(model: dict, gt_volume: Tensor, extra_cond_in=None):
    """evaluate_INN_forward with zero conditions, CWFA.py:134-196 (without the check_empty_depths noise).
    Returns (losses, gt_cache, prior_errors, log_jacobians)."""
    losses, prior, ljs = [], [], []
    cache = [gt_volume]
    x = gt_volume
    for n, lv in enumerate(model["levels"]):
        ch = x.shape[1] // 2
        zeros = torch.zeros((x.shape[0], ch) + tuple(x.shape[2:]), dtype=x.dtype)
        mv = zeros if extra_cond_in is None else extra_cond_in[n]
        z, lo, jac = level_forward(lv["inn"], lv["spec"], x, zeros, mv)
        err = torch.norm(z) ** 2
        losses.append(((0.5 * err - jac) / lo.numel()).mean())
        prior.append(0.5 * err.mean() / lo.numel())
        ljs.append(jac.mean() / lo.numel())
        x = lo
        cache.append(lo)
    return losses, cache, prior, ljs


# ----------------------------------------------------------------------------------------
# Training step of one flow level (CWFA.py:928-1015) -- differentiable through torch CPU autograd
# ----------------------------------------------------------------------------------------
def level_train_loss(inn_sd: SD, cond_sd: SD, spec: dict, gt: Tensor, views: Tensor, mean_vol: Tensor,
                     vol_in: Tensor, cond_weight: float = 0.40984):
    """Loss the reference back-propagates for a flow level (``loss_func_reg='L2'``, z = 0):
    cond = cond_net(views) (CWFA.py:895); vol = inn([0, vol_in], c, rev=True) (:912);
    loss_cond = mse(gt, vol) (:953); Z, J = inn(gt, c) (:966);
    nll = (0.5 * ||Z||^2 - J.mean()) / vol.numel() (:970,978);
    loss = w * loss_cond + (1 - w) * nll (:957,986).  Returns (loss, mse, nll)."""
    c_lf = cond_network(cond_sd, views)
    z0 = torch.zeros_like(vol_in)
    vol, _ = level_inverse(inn_sd, spec, z0, vol_in, c_lf, mean_vol)
    mse = F.mse_loss(gt, vol)
    z, _lo, jac = level_forward(inn_sd, spec, gt, c_lf, mean_vol)
    nll = (0.5 * torch.norm(z) ** 2 - jac.mean()) / vol.numel()
    return cond_weight * mse + (1.0 - cond_weight) * nll, mse, nll


def level_train_grads(inn_sd: SD, cond_sd: SD, spec: dict, gt, views, mean_vol, vol_in, cond_weight: float = 0.40984):
    """Gradients of ``level_train_loss`` w.r.t. every floating-point entry of both state_dicts.
    The shared PReLU of the conditioning net (three keys, one tensor: networks.py:209) is tied before differentiating,
    so its gradient is the sum over its three uses, as in the reference."""
    inn = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in inn_sd.items()}
    cond = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in cond_sd.items()}
    p = "subnetworks.0."
    cond[p + "relu.weight"] = cond[p + "conv1.1.weight"]
    cond[p + "conv3d.1.weight"] = cond[p + "conv1.1.weight"]
    loss, mse, nll = level_train_loss(inn, cond, spec, gt, views, mean_vol, vol_in, cond_weight)
    loss.backward()
    zero = lambda v: torch.zeros_like(v)
    g_inn = {k: (v.grad if v.grad is not None else zero(v)) for k, v in inn.items() if v.dtype.is_floating_point}
    g_cond = {k: (v.grad if v.grad is not None else zero(v)) for k, v in cond.items() if v.dtype.is_floating_point}
    return dict(loss=loss.detach(), mse=mse.detach(), nll=nll.detach(), inn=g_inn, cond=g_cond)


def lion_step(p: Tensor, g: Tensor, m: Tensor, lr: float, beta1: float = 0.9, beta2: float = 0.99, wd: float = 0.0):
    """lion_pytorch 0.0.7 ``update_fn`` (requirements.txt:1; not vendored in the reference tree, restated from the published
    algorithm -- parity unpinned): p *= 1 - lr*wd; p -= lr*sign(b1*m + (1-b1)*g); m = b2*m + (1-b2)*g.  Returns (p, m)."""
    p = p * (1.0 - lr * wd)
    upd = torch.sign(m * beta1 + g * (1.0 - beta1))
    p = p - lr * upd
    m = m * beta2 + g * (1.0 - beta2)
    return p, m


def lrnn_train_grads(sd: SD, views: Tensor, gt: Tensor, mean_vol: Optional[Tensor] = None, bn_mode: str = "batch"):
    """The LRNN ("last step") training loss of the reference, ``loss_func_first_step='L2'``: F.mse_loss(curr_gt, LRNN(views))
    (CWFA.py:882,936-941), differentiated w.r.t. every floating-point entry of the Encoder state_dict (torch CPU autograd)."""
    p = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v) for k, v in sd.items()}
    loss = F.mse_loss(gt, lrnn(p, views, mean_vol, bn_mode))
    loss.backward()
    grads = {k: v.grad for k, v in p.items() if torch.is_tensor(v) and v.requires_grad and v.grad is not None}
    return dict(loss=loss.detach(), grads=grads)


# ----------------------------------------------------------------------------------------
# ActNorm and AllInOneBlock
# ----------------------------------------------------------------------------------------
def actnorm_init(data: Tensor) -> Tuple[Tensor, Tensor]:
    """ActNorm._initialize_with_data, invertible_resnet.py:53-64: scale = log(1/std_c) (unbiased std over batch and space),
    bias = -mean_c(data * exp(scale)).  Returns (scale, bias) shaped (1,C,1,1)."""
    C = data.shape[1]
    flat = data.transpose(0, 1).contiguous().view(C, -1)
    scale = torch.log(1.0 / flat.std(dim=-1))
    bias = -(flat * scale.exp()[:, None]).mean(dim=-1)
    shape = [1, C] + [1] * (data.dim() - 2)
    return scale.view(shape), bias.view(shape)


def actnorm(x: Tensor, scale: Tensor, bias: Tensor, rev: bool = False) -> Tuple[Tensor, Tensor]:
    """ActNorm.forward, invertible_resnet.py:66-81: y = x * exp(scale) + bias, J = sum(scale) * prod(spatial dims) per sample."""
    jac = (scale.sum() * x[0, 0].numel()).repeat(x.shape[0])
    if rev:
        return (x - bias) / scale.exp(), -jac
    return x * scale.exp() + bias, jac


def all_in_one_block(sd: SD, pre: str, x: Tensor, conds: Sequence[Tensor], rev: bool, clamp: float = 2.0, gin: bool = False,
                     global_affine_type: str = "SOFTPLUS", reverse_permutation: bool = False, householder: int = 0):
    """AllInOneBlock.forward, all_in_one_block.py:216-262 (``_permute`` :171-187, ``_pre_permute`` :189-195, ``_affine`` :197-214,
    Householder product :160-169) for image-shaped inputs, with the CWFA sub-network under ``pre + 'subnet.'``."""
    C = x.shape[1]
    if householder:
        w = sd[pre + "w_0"]
        for vk in sd[pre + "vk_householder"]:
            w = torch.mm(w, torch.eye(C, dtype=w.dtype) - 2 * torch.ger(vk, vk) / torch.dot(vk, vk))
        w_perm = w.reshape(C, C, 1, 1)
        w_perm_inv = w_perm.transpose(0, 1).contiguous()
    else:
        w_perm, w_perm_inv = sd[pre + "w_perm"], sd[pre + "w_perm_inv"]
    gs, go = sd[pre + "global_scale"], sd[pre + "global_offset"]
    act = {"SIGMOID": lambda a: 10 * torch.sigmoid(a - 2.0), "SOFTPLUS": lambda a: 0.1 * F.softplus(a, beta=0.5),
           "EXP": torch.exp}[global_affine_type]
    scale = None if gin else act(gs)
    perm_jac = 0.0 if gin else torch.sum(torch.log(scale))
    n_pix = x[0, :1].numel()
    if rev:
        x = F.conv2d(x, w_perm_inv) - go
        if scale is not None:
            x = x / scale
    elif reverse_permutation:
        x = F.conv2d(x, w_perm_inv)
    l1 = C - C // 2
    x1, x2 = x[:, :l1], x[:, l1:]
    a = subnet(sd, pre + "subnet.", torch.cat([x1, *conds], 1) if len(conds) else x1, True) * 0.1
    ch = x2.shape[1]
    s = clamp * torch.tanh(a[:, :ch])
    if gin:
        s = s - s.mean(dim=(1, 2, 3), keepdim=True)
    j2 = s.sum(dim=(1, 2, 3))
    if rev:
        x2, j2 = (x2 - a[:, ch:]) * torch.exp(-s), -j2
    else:
        x2 = x2 * torch.exp(s) + a[:, ch:]
    out = torch.cat((x1, x2), 1)
    if not rev:
        out = F.conv2d((out * scale if scale is not None else out) + go, w_perm)
    elif reverse_permutation:
        out = F.conv2d(out, w_perm)
    return out, j2 + (-1) ** int(rev) * n_pix * perm_jac

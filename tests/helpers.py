"""Shared test helpers: build OUR tiny model with the fixture's deterministic weights."""
import numpy as np
import torch

import cwfa_b200
from oracle.weights import deterministic_fill, seeded_randn


def build_tiny_model(fx, device="cpu"):
    """The tiny config of BASELINE.json configs[0] (D=16,S=64,3 steps) with the golden fixture's
    weights: deterministic_fill seeds from fx['config'], permutations + PermuteDim axes from fx."""
    cfg = fx["config"]
    D, S, MAX = cfg["D"], cfg["S"], cfg["MAX"]
    np.random.seed(0)
    torch.manual_seed(0)
    m = cwfa_b200.CWFAModel(n_depths=D, volume_side_size=S, INN_max_down_steps=MAX)
    seeds = cfg["seeds"]
    for n in range(MAX - 1):
        sd = deterministic_fill(m.conv_inn[n].state_dict(), seeds["inn"] + n)
        sd.update({k: v.clone() for k, v in fx["perms"][n].items()})
        m.conv_inn[n].load_state_dict(sd)
        for node in fx["specs"][n]["nodes"]:
            if node["type"] == "perm_dim":
                m.conv_inn[n].module_list[node["idx"]].dims_to_permute = [1, node["axis"]]
        m.cond_nets[n].load_state_dict(deterministic_fill(m.cond_nets[n].state_dict(), seeds["cond"] + n))
    m.cond_nets[-1].load_state_dict(deterministic_fill(m.cond_nets[-1].state_dict(), seeds["lrnn"]))
    return m.to(device)


def tiny_inputs(fx):
    cfg = fx["config"]
    D, S, MAX = cfg["D"], cfg["S"], cfg["MAX"]
    views = seeded_randn((1, 29, S, S), cfg["seeds"]["views"])
    mean_vols = [seeded_randn((1, D // 2 ** (n + 1), S, S), cfg["seeds"]["mean"] + n, 0.1) for n in range(MAX - 1)]
    mean_vols.append(seeded_randn((1, D // 2 ** (MAX - 1), S, S), cfg["seeds"]["mean"] + MAX - 1, 0.1))
    return views, mean_vols


# ---------------------------------------------------------------------------------------------
# Half-precision OPERAND emulation of the oracle (what the tensor-core path computes): every convolution of the CPU oracle
# rounds its input, weight and output-cotangent operands to bf16 / fp16 and accumulates exactly (run the oracle in float64),
# forward, data gradient and weight gradient alike.  The distance of THIS run from the plain float64 oracle is the error the
# arithmetic type itself causes on a given network -- the yardstick for the GPU kernels' half-precision tolerances.
# ---------------------------------------------------------------------------------------------
import contextlib

import torch.nn.functional as _F


def _rnd(t, kind):
    return t.to(torch.bfloat16 if kind == "bf16" else torch.float16).to(t.dtype)


class _RoundedConv2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, padding, kind, conv):
        xr, wr = _rnd(x, kind), _rnd(w, kind)
        ctx.save_for_backward(xr, wr)
        ctx.padding, ctx.kind, ctx.has_b = padding, kind, b is not None
        return conv(xr, wr, b, padding=padding)

    @staticmethod
    def backward(ctx, dy):
        xr, wr = ctx.saved_tensors
        dyr = _rnd(dy, ctx.kind)
        dx = torch.nn.grad.conv2d_input(xr.shape, wr, dyr, padding=ctx.padding)
        dw = torch.nn.grad.conv2d_weight(xr, wr.shape, dyr, padding=ctx.padding)
        return dx, dw, (dy.sum((0, 2, 3)) if ctx.has_b else None), None, None, None


class _RoundedConvT2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, stride, kind, convt):
        xr, wr = _rnd(x, kind), _rnd(w, kind)
        ctx.save_for_backward(xr, wr)
        ctx.stride, ctx.kind, ctx.has_b = stride, kind, b is not None
        return convt(xr, wr, b, stride=stride)

    @staticmethod
    def backward(ctx, dy):
        xr, wr = ctx.saved_tensors
        dyr = _rnd(dy, ctx.kind)
        with torch.enable_grad():
            xl, wl = xr.detach().requires_grad_(True), wr.detach().requires_grad_(True)
            y = torch.conv_transpose2d(xl, wl, None, stride=ctx.stride)
            dx, dw = torch.autograd.grad(y, (xl, wl), dyr)
        return dx, dw, (dy.sum((0, 2, 3)) if ctx.has_b else None), None, None, None


@contextlib.contextmanager
def half_operand_emulation(kind: str):
    """Inside: ``torch.nn.functional.conv2d`` / ``conv_transpose2d`` (as the oracle calls them) round their operands to
    ``kind`` ('bf16' | 'fp16').  Only plain stride-1 'same' convolutions and stride-2 transposed convolutions are handled --
    exactly what the oracle of the hot path issues."""
    conv, convt = _F.conv2d, _F.conv_transpose2d

    def conv2d(x, w, b=None, stride=1, padding=0, *a, **k):
        assert stride == 1 and not a and not k, "emulation covers the oracle's plain convolutions only"
        return _RoundedConv2d.apply(x, w, b, padding, kind, conv)

    def conv_transpose2d(x, w, b=None, stride=1, *a, **k):
        assert not a and not k
        return _RoundedConvT2d.apply(x, w, b, stride, kind, convt)

    _F.conv2d, _F.conv_transpose2d = conv2d, conv_transpose2d
    try:
        yield
    finally:
        _F.conv2d, _F.conv_transpose2d = conv, convt


# ---------------------------------------------------------------------------------------------
# round-2 fixtures
# ---------------------------------------------------------------------------------------------
def build_filled_model(D, S, MAX, specs, perms, seeds, device="cpu", fill_lrnn=True, **cfg):
    """A CWFAModel of any size whose weights are ``deterministic_fill`` (the same values the golden generator loaded into the
    reference) and whose permutations / PermuteDim axes come from the fixture."""
    np.random.seed(0)
    torch.manual_seed(0)
    m = cwfa_b200.CWFAModel(n_depths=D, volume_side_size=S, INN_max_down_steps=MAX, **cfg)
    for n in range(MAX - 1):
        sd = deterministic_fill(m.conv_inn[n].state_dict(), seeds["inn"] + n)
        sd.update({k: v.clone() for k, v in perms[n].items()})
        m.conv_inn[n].load_state_dict(sd)
        for node in specs[n]["nodes"]:
            if node["type"] == "perm_dim":
                m.conv_inn[n].module_list[node["idx"]].dims_to_permute = [1, node["axis"]]
        if "cond" in seeds:
            m.cond_nets[n].load_state_dict(deterministic_fill(m.cond_nets[n].state_dict(), seeds["cond"] + n))
    if fill_lrnn:
        m.cond_nets[-1].load_state_dict(deterministic_fill(m.cond_nets[-1].state_dict(), seeds["lrnn"]))
    return m.to(device)


def build_full_model(fx, device="cpu"):
    """The FULL config (96 x 512 x 512, 5 steps) with the weights of tests/golden/r2_full.pt."""
    cfg = fx["config"]
    return build_filled_model(cfg["D"], cfg["S"], cfg["MAX"], fx["specs"], fx["perms"], cfg["seeds"], device)


def full_inputs(fx):
    cfg = fx["config"]
    D, S, MAX = cfg["D"], cfg["S"], cfg["MAX"]
    views = seeded_randn((1, 29, S, S), cfg["seeds"]["views"])
    mean_vols = [seeded_randn((1, D // 2 ** (n + 1), S, S), cfg["seeds"]["mean"] + n, 0.1) for n in range(MAX - 1)]
    return views, mean_vols


def probe_positions(numel: int, n: int, seed: int) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randint(0, numel, (n,), generator=g, dtype=torch.int64)


def probe_errors(t: torch.Tensor, pr: dict):
    """(relative norm error, rel-L2 over the probe's sample positions, max-abs over them) of tensor ``t`` against a probe of the
    reference's tensor (tests/golden/make_golden_r2.py:probe)."""
    assert tuple(t.shape) == tuple(pr["shape"]), (tuple(t.shape), pr["shape"])
    flat = t.detach().reshape(-1)
    pos = probe_positions(flat.numel(), pr["n"], pr["seed"]).to(flat.device)
    got = flat[pos].double().cpu()
    ref = pr["values"].double()
    nrm = float(flat.double().norm())
    return abs(nrm - pr["norm"]) / pr["norm"], float((got - ref).norm() / ref.norm()), float((got - ref).abs().max())

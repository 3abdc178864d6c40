"""Invertible modules of the CWFA path with the FrEIA ``InvertibleModule`` contract.

Contract (FrEIA/modules/base.py:7-112): constructed as
``module(dims_in, dims_c=None, **args)`` with shapes that exclude the batch axis;
``forward(x_or_z: tuple, c: tuple = None, rev=False, jac=True) -> (sequence_of_tensors, jac)``
where ``jac`` is a ``(B,)`` tensor or a python number, ``+J`` forward and ``-J`` in reverse;
``output_dims(input_dims)`` for shape inference.  Inputs are never mutated.

Every ``forward`` launches hand-written CUDA kernels through the C ABI (cwfa_b200/ops.py);
tensors must live on a CUDA device.  ``state_dict`` keys and shapes equal the reference's
(``perm`` / ``perm_inv`` are ``nn.Parameter(LongTensor, requires_grad=False)``;
sub-networks keep the unused ``block_grad_up`` / ``block1|block12`` / ``block7|block72``
weights), so reference checkpoints load unchanged.
"""
from __future__ import annotations

import math
import warnings
from typing import Callable, Iterable, List, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from . import ops, packed


class InvertibleModule(nn.Module):
    """Base class (FrEIA/modules/base.py:7-112)."""

    def __init__(self, dims_in: Iterable[Tuple[int]], dims_c: Iterable[Tuple[int]] = None):
        super().__init__()
        self.dims_in = list(dims_in)
        self.dims_c = [] if dims_c is None else list(dims_c)

    def forward(self, x_or_z, c=None, rev: bool = False, jac: bool = True):
        raise NotImplementedError(f"{self.__class__.__name__} does not provide forward(...) method")

    def log_jacobian(self, *args, **kwargs):
        raise DeprecationWarning("module.log_jacobian(...) is deprecated. module.forward(..., jac=True) "
                                 "returns a tuple (out, jacobian) now.")

    def output_dims(self, input_dims: List[Tuple[int]]) -> List[Tuple[int]]:
        raise NotImplementedError(f"{self.__class__.__name__} does not provide output_dims(...)")


# ---------------------------------------------------------------------------------------------
# Haar transforms
# ---------------------------------------------------------------------------------------------
class HaarTransform1D(InvertibleModule):
    """Depth-wise (channel-axis) Haar DWT: lo -> first half of the channels, hi -> second half.
    Shape is unchanged.  Reference: INN_utils.py:126-174 (``order_by_wavelet`` is accepted and
    ignored there too; ``rebalance`` only enters the reported log-det, :135-140,152)."""

    def __init__(self, dims_in, dims_c=None, order_by_wavelet: bool = False, rebalance: float = 1.0):
        super().__init__(dims_in, dims_c)
        self.fac_fwd = 0.5 * rebalance
        self.fac_rev = 0.5 / rebalance
        self.jac_fwd = (np.log(16.0) + 4 * np.log(self.fac_fwd)) / 4.0
        self.jac_rev = (np.log(16.0) + 4 * np.log(self.fac_rev)) / 4.0

    def forward(self, x_in, c=None, jac=True, rev=False):
        x = x_in[0]
        ndims = x[0].numel()
        if not rev:
            return (ops.haar1d_forward(x),), ndims * self.jac_fwd
        return (ops.haar1d_inverse(x),), -ndims * self.jac_rev

    def output_dims(self, input_dims):
        if len(input_dims) != 1:
            raise ValueError("HaarDownsampling must have exactly 1 input")
        if len(input_dims[0]) != 3:
            raise ValueError("HaarDownsampling can only transform 2D images of the shape CxWxH (channels, width, height)")
        c, w, h = input_dims[0]
        return ((c, w, h),)


class HaarDownsampling(InvertibleModule):
    """FrEIA 2-D Haar: (C,H,W) -> (4C,H/2,W/2).  FrEIA/modules/reshapes.py:191-316.
    Unlike the reference (reshapes.py:297) the reverse pass does not scale its input in place."""

    def __init__(self, dims_in, dims_c=None, order_by_wavelet: bool = False, rebalance: float = 1.0):
        super().__init__(dims_in, dims_c)
        if rebalance == 0:
            raise ValueError("'rebalance' argument must be != 0.")
        self.in_channels = dims_in[0][0]
        self.fac_fwd = 0.5 * rebalance
        self.jac_fwd = (np.log(16.0) + 4 * np.log(self.fac_fwd)) / 4.0
        self.fac_rev = 0.5 / rebalance
        self.jac_rev = (np.log(16.0) + 4 * np.log(self.fac_rev)) / 4.0
        # kept for state_dict parity with the reference (a frozen nn.Parameter there, :240-253)
        w = torch.ones(4, 1, 2, 2)
        w[1, 0, 0, 1] = w[1, 0, 1, 1] = -1
        w[2, 0, 1, 0] = w[2, 0, 1, 1] = -1
        w[3, 0, 1, 0] = w[3, 0, 0, 1] = -1
        self.haar_weights = nn.Parameter(torch.cat([w] * self.in_channels, 0), requires_grad=False)
        self.permute = order_by_wavelet

    def forward(self, x, c=None, jac=True, rev=False):
        inp = x[0]
        ndims = inp[0].numel()
        if not rev:
            return (ops.haar2d_down(inp, self.permute, self.fac_fwd),), ndims * self.jac_fwd
        return (ops.haar2d_up(inp, self.permute, self.fac_rev),), ndims * self.jac_rev

    def output_dims(self, input_dims):
        if len(input_dims) != 1:
            raise ValueError("HaarDownsampling must have exactly 1 input")
        if len(input_dims[0]) != 3:
            raise ValueError("HaarDownsampling can only transform 2D images of the shape CxWxH (channels, width, height)")
        c, w, h = input_dims[0]
        c2, w2, h2 = c * 4, w // 2, h // 2
        if c * h * w != c2 * h2 * w2:
            raise ValueError("Input cannot be cleanly reshaped, most likely because the input height or width are an odd number")
        return ((c2, w2, h2),)


class HaarUpsampling(HaarDownsampling):
    """Inverse of HaarDownsampling: (4C,H,W) -> (C,2H,2W).  reshapes.py:319-374."""

    def __init__(self, dims_in, dims_c=None, order_by_wavelet: bool = False, rebalance: float = 1.0):
        inv_shape = self.output_dims(dims_in)
        super().__init__(inv_shape, dims_c, order_by_wavelet, rebalance)

    def forward(self, x, c=None, jac=True, rev=False):
        return super().forward(x, c=None, rev=not rev)

    def output_dims(self, input_dims):
        if len(input_dims) != 1:
            raise ValueError("i-revnet downsampling must have exactly 1 input")
        if len(input_dims[0]) != 3:
            raise ValueError("i-revnet downsampling can only tranform 2d images of the shape cxwxh (channels, width, height)")
        c, w, h = input_dims[0]
        c2, w2, h2 = c // 4, w * 2, h * 2
        if c * h * w != c2 * h2 * w2:
            raise ValueError("input cannot be cleanly reshaped, most likely because the input height or width are an odd number")
        return ((c2, w2, h2),)


# ---------------------------------------------------------------------------------------------
# Split / Concat
# ---------------------------------------------------------------------------------------------
class Split(InvertibleModule):
    """Invertible split along one non-batch axis (FrEIA/modules/graph_topology.py:10-89).
    Forward returns views; reverse concatenates."""

    def __init__(self, dims_in, section_sizes=None, n_sections: int = 2, dim: int = 0):
        super().__init__(dims_in)
        assert len(dims_in) == 1, "Split layer takes exactly one input tensor"
        assert len(dims_in[0]) >= dim, "Split dimension index out of range"
        self.dim = dim
        l_dim = dims_in[0][dim]
        if section_sizes is None:
            assert 2 <= n_sections, "'n_sections' must be a least 2"
            if l_dim % n_sections != 0:
                warnings.warn("Split will create sections of unequal size")
            self.split_size_or_sections = ([l_dim // n_sections + 1] * (l_dim % n_sections)
                                           + [l_dim // n_sections] * (n_sections - l_dim % n_sections))
        else:
            if isinstance(section_sizes, int):
                assert section_sizes < l_dim, "'section_sizes' too large"
            else:
                assert isinstance(section_sizes, (list, tuple)), "'section_sizes' must be either int or list/tuple of int"
                assert sum(section_sizes) <= l_dim, "'section_sizes' too large"
                if sum(section_sizes) < l_dim:
                    warnings.warn("'section_sizes' too small, adding additional section")
                    section_sizes = list(section_sizes) + [l_dim - sum(section_sizes)]
            self.split_size_or_sections = section_sizes

    def forward(self, x, rev=False, jac=True):
        if rev:
            return [torch.cat(list(x), dim=self.dim + 1)], 0
        return torch.split(x[0], self.split_size_or_sections, dim=self.dim + 1), 0

    def output_dims(self, input_dims):
        assert len(input_dims) == 1, "Split layer takes exactly one input tensor"
        sizes = self.split_size_or_sections
        if isinstance(sizes, int):
            l_dim = input_dims[0][self.dim]
            sizes = [sizes] * (l_dim // sizes) + ([l_dim % sizes] if l_dim % sizes else [])
        return [tuple(s if j == self.dim else input_dims[0][j] for j in range(len(input_dims[0]))) for s in sizes]


class Concat(InvertibleModule):
    """Invertible concatenation (FrEIA/modules/graph_topology.py:92-152)."""

    def __init__(self, dims_in, dim: int = 0):
        super().__init__(dims_in)
        assert len(dims_in) > 1, "Concatenation only makes sense for multiple inputs"
        assert len(dims_in[0]) >= dim, "Merge dimension index out of range"
        assert all(len(dims_in[i]) == len(dims_in[0]) for i in range(len(dims_in))), \
            "All input tensors must have same number of dimensions"
        assert all(dims_in[i][j] == dims_in[0][j] for i in range(len(dims_in))
                   for j in range(len(dims_in[i])) if j != dim), \
            "All input tensor dimensions except merge dimension must be identical"
        self.dim = dim
        self.split_size_or_sections = [dims_in[i][dim] for i in range(len(dims_in))]

    def forward(self, x, rev=False, jac=True):
        if rev:
            return torch.split(x[0], self.split_size_or_sections, dim=self.dim + 1), 0
        return [torch.cat(list(x), dim=self.dim + 1)], 0

    def output_dims(self, input_dims):
        assert len(input_dims) > 1, "Concatenation only makes sense for multiple inputs"
        out = list(input_dims[0])
        out[self.dim] = sum(d[self.dim] for d in input_dims)
        return [tuple(out)]


# ---------------------------------------------------------------------------------------------
# Permutations
# ---------------------------------------------------------------------------------------------
def _make_perm(n: int):
    perm = np.random.permutation(n)
    inv = np.zeros_like(perm)
    inv[perm] = np.arange(n)
    return (nn.Parameter(torch.LongTensor(perm), requires_grad=False),
            nn.Parameter(torch.LongTensor(inv), requires_grad=False))


class PermuteRandom(InvertibleModule):
    """Fixed random channel permutation: y = x[:, perm].  fixed_transforms.py:11-46.
    numpy's global RNG is (re)seeded exactly as the reference does so that the same ``seed``
    yields the same permutation."""

    def __init__(self, dims_in, dims_c=None, seed: Union[int, None] = None):
        super().__init__(dims_in, dims_c)
        self.in_channels = dims_in[0][0]
        if seed is not None:
            np.random.seed(seed)
        self.perm, self.perm_inv = _make_perm(self.in_channels)

    def forward(self, x, rev=False, jac=True):
        return [ops.permute(x[0], self.perm_inv if rev else self.perm, 1)], 0.0

    def output_dims(self, input_dims):
        if len(input_dims) != 1:
            raise ValueError(f"{self.__class__.__name__} can only use 1 input")
        return input_dims


class PermuteDim(InvertibleModule):
    """Fixed random permutation of rows (tensor dim 2) or columns (dim 3).  INN_utils.py:46-87.
    As in the reference the axis is drawn from numpy's RNG BEFORE seeding (:61-64) and is not
    part of ``state_dict``; pass ``axis=2|3`` to pin it (e.g. when loading a checkpoint)."""

    def __init__(self, dims_in, dims_c=None, dims_to_permute=[1, 2], seed: Union[int, None] = None,
                 axis: Union[int, None] = None):
        super().__init__(dims_in, dims_c)
        choices = [[1, 2], [1, 3]]
        self.in_channels = dims_in[0][0]
        drawn = choices[np.random.randint(0, len(choices))]
        self.dims_to_permute = drawn if axis is None else [1, int(axis)]
        if seed is not None:
            np.random.seed(seed)
        self.perm, self.perm_inv = _make_perm(dims_in[0][self.dims_to_permute[1] - 1])

    @property
    def axis(self) -> int:
        return int(self.dims_to_permute[1])

    def forward(self, x, rev=False, jac=True):
        return [ops.permute(x[0], self.perm_inv if rev else self.perm, self.axis)], 0.0

    def output_dims(self, input_dims):
        if len(input_dims) != 1:
            raise ValueError(f"{self.__class__.__name__} can only use 1 input")
        return input_dims


# ---------------------------------------------------------------------------------------------
# Coupling blocks
# ---------------------------------------------------------------------------------------------
class _BaseCouplingBlock(InvertibleModule):
    """Dimension checks, split sizes and the soft clamp (coupling_layers.py:8-121).

    Clamp activations (coupling_layers.py:50-60): ``"ATAN"`` (s = clamp * 0.636 * atan(a), what CWFA uses), ``"TANH"``
    (s = clamp * tanh(a)) and ``"SIGMOID"`` (s = clamp * 2 (sigmoid(a) - 0.5) = clamp * tanh(a / 2)) are evaluated inside the
    fused affine kernel (ATAN / TANH clamp modes).  A callable is applied to the sub-network output with the caller's own
    torch ops (it is arbitrary Python) and the kernel receives the finished ``s``."""

    def __init__(self, dims_in, dims_c=[], clamp: float = 2.0, clamp_activation: Union[str, Callable] = "ATAN"):
        super().__init__(dims_in, dims_c)
        self.channels = dims_in[0][0]
        self.ndims = len(dims_in[0])
        self.split_len1 = self.channels // 2
        self.split_len2 = self.channels - self.channels // 2
        self.clamp = clamp
        assert all(tuple(dims_c[i][1:]) == tuple(dims_in[0][1:]) for i in range(len(dims_c))), \
            "Dimensions of input and one or more conditions don't agree."
        self.conditional = len(dims_c) > 0
        self.condition_length = sum(dims_c[i][0] for i in range(len(dims_c)))
        self.f_clamp = None
        if isinstance(clamp_activation, str):
            if clamp_activation == "ATAN":
                self._clamp_kw = dict(k_atan=ops.K_ATAN, tanh_clamp=False)
            elif clamp_activation == "TANH":
                self._clamp_kw = dict(k_atan=1.0, tanh_clamp=True)
            elif clamp_activation == "SIGMOID":
                self._clamp_kw = dict(k_atan=0.5, tanh_clamp=True)          # 2 (sigmoid(u) - 0.5) == tanh(u / 2)
            else:
                raise ValueError(f'Unknown clamp activation "{clamp_activation}"')
        else:
            self.f_clamp = clamp_activation
            self._clamp_kw = None
        self.clamp_activation = clamp_activation

    def _affine(self, x, a_s, a_t, rev, t_scale: float = 1.0):
        """y, log-det of one affine coupling with this block's clamp: the fused affine kernel."""
        if self._clamp_kw is None:                                     # user-supplied clamp function
            s = self.clamp * self.f_clamp(a_s)
            return ops.affine(x, s.contiguous(), a_t, inverse=rev, clamp=self.clamp, t_scale=t_scale, s_is_final=True)
        return ops.affine(x, a_s, a_t, inverse=rev, clamp=self.clamp, t_scale=t_scale, **self._clamp_kw)

    def _fused_executor(self, subnet, n_out: int, ext: bool, *tensors):
        """The sub-network's tensor-core executor when this coupling can run with the coupling FUSED into the epilogue of the
        sub-network's last convolution (``tc.conv_tc_coupling``), else None: needs the inference switch on
        (``cwfa_b200.set_inference_precision``), no gradient, the ATAN clamp (the only one the epilogue evaluates), image-shaped
        data and [s | t] (or s alone when the shift is external) in one N block of the last conv."""
        kind = packed.fast_kind(*tensors)
        ex_fn = getattr(subnet, "packed_executor", None)
        if kind is None or ex_fn is None or self._clamp_kw is None or self._clamp_kw["tanh_clamp"] or self.ndims != 3 or n_out > 48:
            return None
        ex = ex_fn(kind)
        return ex if ex.out.Cout == (n_out if ext else 2 * n_out) and ex.out.BN == ex.out.Cout_p else None

    def _couple(self, subnet, u, x_active, n_out: int, rev: bool):
        """(y, log-det) of one affine coupling whose coefficients are ``subnet(u)`` = [s_raw | t]: ONE tensor-core kernel for
        the last conv + coupling when ``_fused_executor`` allows it (s, t never reach HBM), else sub-network output -> fused
        affine kernel."""
        ex = self._fused_executor(subnet, n_out, False, u, x_active)
        if ex is not None:
            return ex.couple(u, x_active, ch=n_out, inverse=rev, clamp=self.clamp, k_atan=self._clamp_kw["k_atan"])
        a = subnet(u)
        return self._affine(x_active, a[:, :n_out], a[:, n_out:], rev)

    def _clamped_s(self, a_s):
        """s = clamp * f_clamp(a_s) as a tensor (only the volume-preserving GIN block needs it outside the kernel)."""
        if self._clamp_kw is None:
            return self.clamp * self.f_clamp(a_s)
        k = self._clamp_kw["k_atan"]
        return self.clamp * (torch.tanh(k * a_s) if self._clamp_kw["tanh_clamp"] else k * torch.atan(a_s))

    def output_dims(self, input_dims):
        if len(input_dims) != 1:
            raise ValueError("Can only use 1 input")
        return input_dims

    # ---- shared two-sided driver (coupling_layers.py:62-87) ----
    def forward(self, x, c=[], rev=False, jac=True):
        x1, x2 = torch.split(x[0], [self.split_len1, self.split_len2], dim=1)
        c = list(c) if c is not None else []
        if not rev:
            x2_c = torch.cat([x2, *c], 1) if self.conditional else x2
            y1, j1 = self._coupling1(x1, x2_c)
            y1_c = torch.cat([y1, *c], 1) if self.conditional else y1
            y2, j2 = self._coupling2(x2, y1_c)
        else:
            x1_c = torch.cat([x1, *c], 1) if self.conditional else x1
            y2, j2 = self._coupling2(x2, x1_c, rev=True)
            y2_c = torch.cat([y2, *c], 1) if self.conditional else y2
            y1, j1 = self._coupling1(x1, y2_c, rev=True)
        return (torch.cat((y1, y2), 1),), j1 + j2

    def _affine_from(self, x_active, a, n_out, rev, gin=False):
        s_raw, t = a[:, :n_out], a[:, n_out:]
        if gin:
            # volume preserving: s -= mean over channels (coupling_layers.py:361)
            s = self._clamped_s(s_raw)
            s = s - s.mean(1, keepdim=True)
            y, _ = ops.affine(x_active, s.contiguous(), t, inverse=rev, clamp=self.clamp, s_is_final=True)
            return y, 0.0
        return self._affine(x_active, s_raw, t, rev)


class NICECouplingBlock(_BaseCouplingBlock):
    """Additive coupling (coupling_layers.py:124-157): the affine kernel with s = 0."""

    def __init__(self, dims_in, dims_c=[], subnet_constructor: Callable = None):
        super().__init__(dims_in, dims_c, clamp=0.0)
        self.F = subnet_constructor(self.split_len2 + self.condition_length, self.split_len1)
        self.G = subnet_constructor(self.split_len1 + self.condition_length, self.split_len2)

    def _coupling1(self, x1, u2, rev=False):
        t = self.F(u2)
        y, _ = ops.affine(x1, t, t, inverse=rev, clamp=0.0)
        return y, 0.0

    def _coupling2(self, x2, u1, rev=False):
        t = self.G(u1)
        y, _ = ops.affine(x2, t, t, inverse=rev, clamp=0.0)
        return y, 0.0


class RNVPCouplingBlock(_BaseCouplingBlock):
    """RealNVP-style block with four sub-networks (coupling_layers.py:160-229)."""

    def __init__(self, dims_in, dims_c=[], subnet_constructor: Callable = None, clamp: float = 2.0,
                 clamp_activation: Union[str, Callable] = "ATAN"):
        super().__init__(dims_in, dims_c, clamp, clamp_activation)
        self.subnet_s1 = subnet_constructor(self.split_len1 + self.condition_length, self.split_len2)
        self.subnet_t1 = subnet_constructor(self.split_len1 + self.condition_length, self.split_len2)
        self.subnet_s2 = subnet_constructor(self.split_len2 + self.condition_length, self.split_len1)
        self.subnet_t2 = subnet_constructor(self.split_len2 + self.condition_length, self.split_len1)

    def _st_couple(self, sub_s, sub_t, u, x_active, n_out, rev):
        t = sub_t(u)
        ex = self._fused_executor(sub_s, n_out, True, u, x_active)
        if ex is not None:                       # s from the fused last conv of subnet_s, t (its own sub-network) as external shift
            return ex.couple(u, x_active, ch=n_out, inverse=rev, clamp=self.clamp, k_atan=self._clamp_kw["k_atan"], t_ext=t, t_scale=1.0)
        return self._affine(x_active, sub_s(u), t, rev)

    def _coupling1(self, x1, u2, rev=False):
        return self._st_couple(self.subnet_s2, self.subnet_t2, u2, x1, self.split_len1, rev)

    def _coupling2(self, x2, u1, rev=False):
        return self._st_couple(self.subnet_s1, self.subnet_t1, u1, x2, self.split_len2, rev)


class GLOWCouplingBlock(_BaseCouplingBlock):
    """GLOW-style block: one sub-network predicts [s, t] jointly (coupling_layers.py:232-302)."""

    def __init__(self, dims_in, dims_c=[], subnet_constructor: Callable = None, clamp: float = 2.0,
                 clamp_activation: Union[str, Callable] = "ATAN"):
        super().__init__(dims_in, dims_c, clamp, clamp_activation)
        self.subnet1 = subnet_constructor(self.split_len1 + self.condition_length, self.split_len2 * 2)
        self.subnet2 = subnet_constructor(self.split_len2 + self.condition_length, self.split_len1 * 2)
        self._gin = False

    def _coupling1(self, x1, u2, rev=False):
        if self._gin:
            return self._affine_from(x1, self.subnet2(u2), self.split_len1, rev, True)
        return self._couple(self.subnet2, u2, x1, self.split_len1, rev)

    def _coupling2(self, x2, u1, rev=False):
        if self._gin:
            return self._affine_from(x2, self.subnet1(u1), self.split_len2, rev, True)
        return self._couple(self.subnet1, u1, x2, self.split_len2, rev)


class GINCouplingBlock(GLOWCouplingBlock):
    """Volume-preserving GLOW variant (coupling_layers.py:305-381)."""

    def __init__(self, dims_in, dims_c=[], subnet_constructor: Callable = None, clamp: float = 2.0,
                 clamp_activation: Union[str, Callable] = "ATAN"):
        super().__init__(dims_in, dims_c, subnet_constructor, clamp, clamp_activation)
        self._gin = True


class AffineCouplingOneSided(_BaseCouplingBlock):
    """Half of a GLOW block (coupling_layers.py:384-437)."""

    def __init__(self, dims_in, dims_c=[], subnet_constructor: Callable = None, clamp: float = 2.0,
                 clamp_activation: Union[str, Callable] = "ATAN"):
        super().__init__(dims_in, dims_c, clamp, clamp_activation)
        self.subnet = subnet_constructor(self.split_len1 + self.condition_length, 2 * self.split_len2)

    def forward(self, x, c=[], rev=False, jac=True):
        x1, x2 = torch.split(x[0], [self.split_len1, self.split_len2], dim=1)
        x1_c = torch.cat([x1, *c], 1) if self.conditional else x1
        y2, j = self._couple(self.subnet, x1_c, x2, self.split_len2, rev)
        return (torch.cat((x1, y2), 1),), j


class ConditionalAffineTransform(_BaseCouplingBlock):
    """CWFA's default block ("CAT"): s, t predicted from the CONDITION only and applied to the
    whole input (coupling_layers.py:440-500).  ``y = exp(s) x + t`` / ``y = (x - t) exp(-s)``,
    ``s = clamp * 0.636 * atan(a[:, :ch])``, log-det = +-sum(s) per sample -- one fused kernel.

    When the sub-network is CWFA's ``_first`` variant (t = -meanvol/sqrt2, networks.py:653-671)
    the shift half is never materialised: the kernel reads the condition with t_scale=-1/sqrt2.
    """

    def __init__(self, dims_in, dims_c=[], subnet_constructor: Callable = None, clamp: float = 2.0,
                 clamp_activation: Union[str, Callable] = "ATAN"):
        super().__init__(dims_in, dims_c, clamp, clamp_activation)
        if not self.conditional:
            raise ValueError("ConditionalAffineTransform must have a condition")
        self.subnet = subnet_constructor(self.condition_length, 2 * self.channels)

    def _st(self, c):
        cond = torch.cat(list(c), 1) if len(c) > 1 else c[0]
        split_fn = getattr(self.subnet, "forward_split", None)
        if split_fn is not None:
            return split_fn(cond)
        a = self.subnet(cond)
        return a[:, :self.channels], a[:, self.channels:], 1.0

    def forward(self, x, c=[], rev=False, jac=True):
        memo = _SUBNET_MEMO[-1] if _SUBNET_MEMO else None
        first = not getattr(self.subnet, "normal", True)
        ex = self._fused_executor(self.subnet, self.channels, first, x[0], *c) if memo is None else None
        if ex is not None:
            # tensor-core inference: sub-network trunk + last conv with the coupling in its epilogue
            cond = torch.cat(list(c), 1) if len(c) > 1 else c[0]
            if first:                                                    # s = trunk(LF half), t = -meanvol / sqrt2 (networks.py:653-671)
                n = self.subnet.c_in // 2
                y, j = ex.couple(cond[:, -n:].contiguous(), x[0], ch=self.channels, inverse=rev, clamp=self.clamp,
                                 k_atan=self._clamp_kw["k_atan"], t_ext=cond[:, :-n].contiguous(), t_scale=-1.0 / math.sqrt(2))
            else:
                y, j = ex.couple(cond, x[0], ch=self.channels, inverse=rev, clamp=self.clamp, k_atan=self._clamp_kw["k_atan"])
            return (y,), j
        if memo is None:
            a_s, a_t, t_scale = self._st(c)
        else:
            # (s, t) depend on the conditions only: a training step that runs the level in both directions on the SAME
            # condition tensors (CWFA.py:912 and :966) evaluates each sub-network once; autograd sums both uses.
            key = (id(self),) + tuple((id(t), t._version) for t in c)
            hit = memo.get(key)
            if hit is None:
                hit = memo[key] = (self._st(c), list(c))          # keeps the condition tensors alive: ids stay unique
            a_s, a_t, t_scale = hit[0]
        y, j = self._affine(x[0], a_s, a_t, rev, t_scale=t_scale)
        return (y,), j


_SUBNET_MEMO: list = []


class share_subnet_outputs:
    """Context manager: inside it a ``ConditionalAffineTransform`` called again with the same condition tensor objects
    reuses its sub-network output instead of recomputing it (used by ``cwfa_b200.training.flow_level_loss``)."""

    def __enter__(self):
        _SUBNET_MEMO.append({})
        return self

    def __exit__(self, *exc):
        _SUBNET_MEMO.pop()
        return False


# ---------------------------------------------------------------------------------------------
# ActNorm and AllInOneBlock (the remaining selectable module-API surface)
# ---------------------------------------------------------------------------------------------
class ActNorm(InvertibleModule):
    """Per-channel affine y = x * exp(scale) + bias, initialised from the first batch to zero mean / unit std
    (FrEIA/modules/invertible_resnet.py:11-85).  One ``scale_shift`` kernel per call; the data-dependent initialisation
    reads the per-channel (sum, sum of squares) from the ``channel_stats`` kernel."""

    def __init__(self, dims_in, dims_c=None, init_data: Union[torch.Tensor, None] = None):
        super().__init__(dims_in, dims_c)
        self.dims_in = dims_in[0]
        param_dims = [1, self.dims_in[0]] + [1 for _ in range(len(self.dims_in) - 1)]
        self.scale = nn.Parameter(torch.zeros(*param_dims))
        self.bias = nn.Parameter(torch.zeros(*param_dims))
        self.init_on_next_batch = init_data is None
        if init_data is not None:
            self._initialize_with_data(init_data)
        self._register_load_state_dict_pre_hook(lambda *a: setattr(self, "init_on_next_batch", False))

    @torch.no_grad()
    def _initialize_with_data(self, data):
        assert all(data.shape[i + 1] == self.dims_in[i] for i in range(len(self.dims_in))), \
            "Can't initialize ActNorm layer, provided data don't match input dimensions."
        C = self.dims_in[0]
        x = data.reshape(data.shape[0], C, -1)
        n = x.shape[0] * x.shape[2]
        s, q = ops.channel_stats(x.reshape(x.shape[0], C, 1, -1)).double()
        var = (q - s * s / n) / (n - 1)                                   # torch.std: unbiased
        scale = torch.log(1.0 / var.sqrt())
        self.scale.data = scale.float().view_as(self.scale).to(self.scale.device)
        self.bias.data = (-(s / n) * scale.exp()).float().view_as(self.bias).to(self.bias.device)
        self.init_on_next_batch = False

    def forward(self, x, rev=False, jac=True):
        if self.init_on_next_batch:
            self._initialize_with_data(x[0])
        # the C-element parameter algebra (exp, reciprocal) is host-level glue on torch's tape; the per-element pass and its
        # adjoint (dx, d scale, d bias) are the scale_shift kernels (cwfa_b200.autograd._ScaleShift)
        j = (self.scale.sum() * np.prod(self.dims_in[1:])).repeat(x[0].shape[0])
        e = self.scale.reshape(-1).exp()
        b = self.bias.reshape(-1)
        if not rev:
            return [ops.scale_shift(x[0], e, b)], j
        return [ops.scale_shift(x[0], 1.0 / e, -b / e)], -j

    def output_dims(self, input_dims):
        assert len(input_dims) == 1, "Can only use 1 input"
        return input_dims


class AllInOneBlock(InvertibleModule):
    """Coupling + (soft / hard / Householder) permutation + global affine in one block
    (FrEIA/modules/all_in_one_block.py:13-271), for image-shaped inputs (rank 2).
    ``y = W (Psi(s_global) * Coupling(x) + t_global)``; the coupling is ``x2 * exp(clamp * tanh(0.1 a_s)) + 0.1 a_t`` with
    ``a = subnet(cat(x1, c))``.  The coupling runs in the fused affine kernel (TANH clamp mode), the channel mixing as a
    1x1 convolution, the global affine as one scale/shift pass."""

    def __init__(self, dims_in, dims_c=[], subnet_constructor: Callable = None, affine_clamping: float = 2.0,
                 gin_block: bool = False, global_affine_init: float = 1.0, global_affine_type: str = "SOFTPLUS",
                 permute_soft: bool = False, learned_householder_permutation: int = 0, reverse_permutation: bool = False):
        super().__init__(dims_in, dims_c)
        channels = dims_in[0][0]
        self.input_rank = len(dims_in[0]) - 1
        if self.input_rank != 2:
            raise ValueError("cwfa_b200.AllInOneBlock implements image-shaped inputs (C,H,W) only")
        self.sum_dims = tuple(range(1, 2 + self.input_rank))
        if len(dims_c) == 0:
            self.conditional, self.condition_channels = False, 0
        else:
            assert tuple(dims_c[0][1:]) == tuple(dims_in[0][1:]), \
                f"Dimensions of input and condition don't agree: {dims_c} vs {dims_in}."
            self.conditional, self.condition_channels = True, sum(dc[0] for dc in dims_c)
        self.splits = [channels - channels // 2, channels // 2]
        self.in_channels, self.clamp, self.GIN = channels, affine_clamping, gin_block
        self.reverse_pre_permute, self.householder = reverse_permutation, learned_householder_permutation
        if permute_soft and channels > 512:
            warnings.warn(f"Soft permutation will take a very long time to initialize with {channels} feature channels. "
                          "Consider using hard permutation instead.")
        if global_affine_type == "SIGMOID":
            global_scale = 2.0 - np.log(10.0 / global_affine_init - 1.0)
            self.global_scale_activation = lambda a: 10 * torch.sigmoid(a - 2.0)
        elif global_affine_type == "SOFTPLUS":
            global_scale = 2.0 * np.log(np.exp(0.5 * 10.0 * global_affine_init) - 1)
            self.softplus = nn.Softplus(beta=0.5)
            self.global_scale_activation = lambda a: 0.1 * self.softplus(a)
        elif global_affine_type == "EXP":
            global_scale = np.log(global_affine_init)
            self.global_scale_activation = lambda a: torch.exp(a)
        else:
            raise ValueError('Global affine activation must be "SIGMOID", "SOFTPLUS" or "EXP"')
        self.global_scale = nn.Parameter(torch.ones(1, channels, 1, 1) * float(global_scale))
        self.global_offset = nn.Parameter(torch.zeros(1, channels, 1, 1))
        if permute_soft:
            from scipy.stats import special_ortho_group
            w = special_ortho_group.rvs(channels)
        else:
            w = np.zeros((channels, channels))
            for i, j in enumerate(np.random.permutation(channels)):
                w[i, j] = 1.0
        if self.householder:
            self.vk_householder = nn.Parameter(0.2 * torch.randn(self.householder, channels), requires_grad=True)
            self.w_perm = self.w_perm_inv = None
            self.w_0 = nn.Parameter(torch.FloatTensor(w), requires_grad=False)
        else:
            self.w_perm = nn.Parameter(torch.FloatTensor(w).view(channels, channels, 1, 1), requires_grad=False)
            self.w_perm_inv = nn.Parameter(torch.FloatTensor(w.T).view(channels, channels, 1, 1), requires_grad=False)
        if subnet_constructor is None:
            raise ValueError("Please supply a callable subnet_constructor function or object (see docstring)")
        self.subnet = subnet_constructor(self.splits[0] + self.condition_channels, 2 * self.splits[1])
        self.last_jac = None

    def _construct_householder_permutation(self):
        w = self.w_0
        for vk in self.vk_householder:                      # C x C host-side algebra on the reflection vectors
            w = torch.mm(w, torch.eye(self.in_channels, device=w.device) - 2 * torch.ger(vk, vk) / torch.dot(vk, vk))
        return w.reshape(self.in_channels, self.in_channels, 1, 1)

    def _permute(self, x, rev=False):
        if self.GIN:
            scale, perm_log_jac = None, 0.0
        else:
            scale = self.global_scale_activation(self.global_scale).reshape(-1)
            perm_log_jac = torch.sum(torch.log(scale))
        off = self.global_offset.reshape(-1)
        # differentiable end to end: scale_shift and the 1x1 channel mixing have adjoint kernels (global_scale, global_offset
        # and -- through the Householder product -- vk_householder receive gradients, as in the reference)
        if rev:
            y = ops.conv2d(x, self.w_perm_inv, None)
            return (ops.scale_shift(y, torch.ones_like(off) if scale is None else 1.0 / scale, -off if scale is None else -off / scale),
                    perm_log_jac)
        y = ops.scale_shift(x, torch.ones_like(off) if scale is None else scale, off)
        return ops.conv2d(y, self.w_perm, None), perm_log_jac

    def _pre_permute(self, x, rev=False):
        return ops.conv2d(x, self.w_perm if rev else self.w_perm_inv, None)

    def _affine(self, x, a, rev=False):
        ch = x.shape[1]
        if self.GIN:
            s = self.clamp * torch.tanh(0.1 * a[:, :ch])
            s = s - torch.mean(s, dim=self.sum_dims, keepdim=True)
            y, _ = ops.affine(x, s.contiguous(), a[:, ch:], inverse=rev, t_scale=0.1, s_is_final=True)
            j = torch.sum(s, dim=self.sum_dims)
            return y, (-j if rev else j)
        return ops.affine(x, a[:, :ch], a[:, ch:], inverse=rev, clamp=self.clamp, k_atan=0.1, t_scale=0.1, tanh_clamp=True)

    def forward(self, x, c=[], rev=False, jac=True):
        if self.householder:
            self.w_perm = self._construct_householder_permutation()
            if rev or self.reverse_pre_permute:
                self.w_perm_inv = self.w_perm.transpose(0, 1).contiguous()
        global_scaling_jac = 0.0
        if rev:
            x0, global_scaling_jac = self._permute(x[0], rev=True)
        elif self.reverse_pre_permute:
            x0 = self._pre_permute(x[0], rev=False)
        else:
            x0 = x[0]
        x1, x2 = torch.split(x0, self.splits, dim=1)
        x1c = torch.cat([x1, *c], 1) if self.conditional else x1.contiguous()
        a1 = self.subnet(x1c)
        x2, j2 = self._affine(x2.contiguous(), a1, rev=rev)
        x_out = torch.cat((x1, x2), 1)
        if not rev:
            x_out, global_scaling_jac = self._permute(x_out, rev=False)
        elif self.reverse_pre_permute:
            x_out = self._pre_permute(x_out, rev=True)
        n_pixels = x_out[0, :1].numel()
        return (x_out,), j2 + (-1) ** rev * n_pixels * global_scaling_jac

    def output_dims(self, input_dims):
        return input_dims

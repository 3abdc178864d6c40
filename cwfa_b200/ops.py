"""Tensor-level wrappers over the C ABI (torch is used only for device memory and streams).

Every function takes contiguous fp32 CUDA tensors in the reference's NCHW layout, launches
on the CURRENT torch stream and returns freshly allocated tensors.  Non-CUDA tensors are
rejected: there is no CPU fallback.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _lib

ACT_NONE, ACT_ELU, ACT_PRELU, ACT_RELU, ACT_GELU, ACT_SIGMOID = range(6)

CLAMP_DEFAULT = 2.0
K_ATAN = 0.636          # FrEIA/modules/coupling_layers.py:52


def _ck(t: torch.Tensor, name: str = "tensor", dtype=torch.float32) -> torch.Tensor:
    if not torch.is_tensor(t):
        raise TypeError(f"{name}: expected a tensor")
    if not t.is_cuda:
        raise RuntimeError(f"cwfa_b200: {name} is on {t.device}; the CUDA kernels are the only "
                           "implementation (no CPU fallback)")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _grad_on(*tensors) -> bool:
    """True when the call must go on the autograd tape (cwfa_b200/autograd.py: forward and adjoint both run this repo's
    kernels).  Inside an autograd Function's forward/backward gradients are disabled, so the raw launches below run."""
    return torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in tensors)


def _no_adjoint(what: str) -> None:
    """An op without an adjoint kernel was reached while an input requires a gradient.  Silently detaching would cut the
    parameters upstream off from the loss (they would then drift under weight decay / partial gradients): refuse."""
    raise NotImplementedError(f"cwfa_b200: {what} has no adjoint kernel; run it under torch.no_grad() or detach its inputs explicitly")


def perm_i32(perm: torch.Tensor, device) -> torch.Tensor:
    """int32 device copy of a LongTensor permutation, cached ON the source tensor object (keyed by
    device and in-place version) so that the cache can never outlive or alias the permutation."""
    cache = getattr(perm, "_cwfa_i32", None)
    key = (str(device), perm._version)
    if cache is None or cache[0] != key:
        cache = (key, perm.detach().to(device=device, dtype=torch.int32).contiguous())
        try:
            perm._cwfa_i32 = cache
        except AttributeError:      # plain tensors without __dict__ support: do not cache
            pass
    return cache[1]


# ---------------------------------------------------------------------------------------------
def haar1d_forward(x: torch.Tensor) -> torch.Tensor:
    """(B,C,H,W) -> (B,C,H,W) with [:, :C/2] = lo, [:, C/2:] = hi.  INN_utils.py:153-156."""
    if _grad_on(x):
        from . import autograd as ag
        return ag.haar1d(x, False)
    x = _ck(x, "x")
    B, C = x.shape[0], x.shape[1]
    P = x[0, 0].numel()
    out = torch.empty_like(x)
    h = C // 2
    _lib.call("cwfa_haar1d_fwd", x.data_ptr(), out.data_ptr(), out.data_ptr() + 4 * h * P,
              B, C, P, C * P, C * P, _stream())
    return out


def haar1d_inverse(x: torch.Tensor) -> torch.Tensor:
    """Inverse of haar1d_forward on one (B,C,H,W) tensor.  INN_utils.py:157-160."""
    if _grad_on(x):
        from . import autograd as ag
        return ag.haar1d(x, True)
    x = _ck(x, "x")
    B, C = x.shape[0], x.shape[1]
    P = x[0, 0].numel()
    out = torch.empty_like(x)
    h = C // 2
    _lib.call("cwfa_haar1d_inv", x.data_ptr(), x.data_ptr() + 4 * h * P, out.data_ptr(),
              B, C, P, C * P, C * P, _stream())
    return out


def haar1d_merge(lo: torch.Tensor, hi: torch.Tensor) -> torch.Tensor:
    """Fused Split^-1 + IDWT: two (B,C/2,H,W) tensors -> (B,C,H,W) without the concat copy."""
    if _grad_on(lo, hi):
        from . import autograd as ag
        return ag.haar1d_merge(lo, hi)
    lo, hi = _ck(lo, "lo"), _ck(hi, "hi")
    B, h = lo.shape[0], lo.shape[1]
    P = lo[0, 0].numel()
    out = torch.empty((B, 2 * h) + tuple(lo.shape[2:]), device=lo.device, dtype=torch.float32)
    _lib.call("cwfa_haar1d_inv", lo.data_ptr(), hi.data_ptr(), out.data_ptr(), B, 2 * h, P, h * P, h * P, _stream())
    return out


def haar1d_split(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fused DWT + Split: (B,C,H,W) -> separate contiguous (lo, hi)."""
    if _grad_on(x):
        from . import autograd as ag
        return ag.haar1d_split(x)
    x = _ck(x, "x")
    B, C = x.shape[0], x.shape[1]
    P = x[0, 0].numel()
    h = C // 2
    lo = torch.empty((B, h) + tuple(x.shape[2:]), device=x.device, dtype=torch.float32)
    hi = torch.empty_like(lo)
    _lib.call("cwfa_haar1d_fwd", x.data_ptr(), lo.data_ptr(), hi.data_ptr(), B, C, P, h * P, h * P, _stream())
    return lo, hi


def haar2d_down(x: torch.Tensor, order_by_wavelet: bool, fac: float) -> torch.Tensor:
    x = _ck(x, "x")
    B, C, H, W = x.shape
    y = torch.empty((B, 4 * C, H // 2, W // 2), device=x.device, dtype=torch.float32)
    _lib.call("cwfa_haar2d_down", x.data_ptr(), y.data_ptr(), B, C, H, W, int(order_by_wavelet), float(fac), _stream())
    return y


def haar2d_up(y: torch.Tensor, order_by_wavelet: bool, fac: float) -> torch.Tensor:
    y = _ck(y, "y")
    B, C4, H2, W2 = y.shape
    C = C4 // 4
    x = torch.empty((B, C, 2 * H2, 2 * W2), device=y.device, dtype=torch.float32)
    _lib.call("cwfa_haar2d_up", y.data_ptr(), x.data_ptr(), B, C, 2 * H2, 2 * W2, int(order_by_wavelet), float(fac), _stream())
    return x


def permute(x: torch.Tensor, perm: torch.Tensor, axis: int) -> torch.Tensor:
    """y = x.index_select(axis, perm) for axis in {1,2,3} of a (B,C,H,W) tensor."""
    if _grad_on(x):
        from . import autograd as ag
        return ag.permute(x, perm, axis)
    x = _ck(x, "x")
    if x.dim() != 4:
        x4 = x.reshape(x.shape[0], x.shape[1], 1, -1)
    else:
        x4 = x
    B, C, H, W = x4.shape
    p = perm_i32(perm, x.device)
    if p.numel() != x4.shape[axis]:
        raise ValueError(f"permutation of length {p.numel()} applied to axis {axis} of size {x4.shape[axis]}")
    y = torch.empty_like(x4)
    _lib.call("cwfa_permute", x4.data_ptr(), y.data_ptr(), p.data_ptr(), axis, B, C, H, W, _stream())
    return y.view(x.shape)


def _prep_inner(t: torch.Tensor):
    """fp32 tensor whose per-sample (ch,H,W) block is contiguous, plus its batch stride in elements."""
    B, ch = t.shape[0], t.shape[1]
    P = t[0, 0].numel()
    if t.dtype != torch.float32:
        t = t.float()
    inner_ok = t[0].is_contiguous() if B > 0 else True
    if B > 1 and inner_ok and t.stride(0) == 0:
        return t, 0                                   # one (ch,H,W) block broadcast over the batch (``expand``): read in place
    if not inner_ok or (B > 1 and t.stride(0) < ch * P):
        t = t.contiguous()
    return t, (t.stride(0) if B > 1 else ch * P)


def affine(x: Optional[torch.Tensor], a_s: torch.Tensor, a_t: torch.Tensor, *, inverse: bool,
           clamp: float = CLAMP_DEFAULT, t_scale: float = 1.0, want_sumsq: bool = False,
           k_atan: float = K_ATAN, s_is_final: bool = False, tanh_clamp: bool = False):
    """Affine coupling with fused log-det.  a_s / a_t: (B,ch,H,W) views whose inner (ch,H,W)
    block is contiguous (e.g. the two channel halves of one subnet output).
    ``tanh_clamp``: s = clamp * tanh(k_atan * a_s) (AllInOneBlock, all_in_one_block.py:206-211) instead of the ATAN clamp."""
    if not a_s.is_cuda:
        raise RuntimeError("cwfa_b200: affine needs CUDA tensors (no CPU fallback)")
    if _grad_on(x, a_s, a_t):
        from . import autograd as ag
        y, logdet = ag.affine(x, a_s, a_t, inverse=inverse, clamp=clamp, t_scale=t_scale, k_atan=k_atan,
                              s_is_final=2 if tanh_clamp else int(bool(s_is_final)))
        return (y, logdet, ag.sum_squares(y)) if want_sumsq else (y, logdet)
    B, ch = a_s.shape[0], a_s.shape[1]
    P = a_s[0, 0].numel()
    a_s, ld_s = _prep_inner(a_s)
    a_t, ld_t = _prep_inner(a_t)
    xx = None if x is None else _ck(x, "x")
    y = torch.empty((B, ch) + tuple(a_s.shape[2:]), device=a_s.device, dtype=torch.float32)
    logdet = torch.empty(B, device=a_s.device, dtype=torch.float32)
    sumsq = torch.empty(B, device=a_s.device, dtype=torch.float32) if want_sumsq else None
    nblk = _lib.load().cwfa_affine_workspace_blocks()
    ws = torch.empty(2 * B * nblk, device=a_s.device, dtype=torch.float32)
    _lib.call("cwfa_affine", _p(xx), a_s.data_ptr(), a_t.data_ptr(), y.data_ptr(), logdet.data_ptr(), _p(sumsq),
              ws.data_ptr(), B, ch, P, ld_s, ld_t, float(clamp), float(k_atan), float(t_scale), int(inverse) | (4 if (tanh_clamp or s_is_final == 2) else (2 if s_is_final else 0)), _stream())
    return (y, logdet, sumsq) if want_sumsq else (y, logdet)


def conv2d(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, act: int = ACT_NONE,
           slope: Optional[torch.Tensor] = None, res: Optional[torch.Tensor] = None, res_mode: int = 0) -> torch.Tensor:
    """fp32 'same' convolution, stride 1.  v = conv+bias; (res_mode 1: +res); act; (res_mode 2: +res)."""
    if _grad_on(x, w, bias, slope, res):
        from . import autograd as ag
        if ag.conv2d_supported(w, act, res, res_mode):
            return ag.conv2d(x, w, bias, act=act, slope=slope, res=res, res_mode=res_mode)
        _no_adjoint(f"conv2d {tuple(w.shape[2:])} act={act} res_mode={res_mode}")
    x, w = _ck(x, "x"), _ck(w, "w")
    N, Cin, H, W = x.shape
    Cout, Cin_w, KH, KW = w.shape
    if Cin_w != Cin:
        raise ValueError(f"conv2d: weight expects {Cin_w} input channels, got {Cin}")
    bias = None if bias is None else _ck(bias, "bias")
    res = None if res is None else _ck(res, "res")
    slope = None if slope is None else _ck(slope, "slope")
    y = torch.empty((N, Cout, H, W), device=x.device, dtype=torch.float32)
    _lib.call("cwfa_conv2d_f32", x.data_ptr(), w.data_ptr(), _p(bias), _p(res), _p(slope), y.data_ptr(),
              N, Cin, H, W, Cout, KH, KW, act, res_mode if res is not None else 0, _stream())
    return y


def conv1d_flat(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], act: int) -> torch.Tensor:
    """Conv1d over the flattened H*W axis (GlobalAttention, networks.py:250-262): (B,C,L)."""
    if _grad_on(x, w, bias):
        from . import autograd as ag
        return ag.conv1d_flat(x, w, bias, act)
    B, C, L = x.shape
    k = w.shape[-1]
    y = conv2d(x.reshape(B, C, 1, L), w.reshape(w.shape[0], w.shape[1], 1, k), bias, act=act)
    return y.reshape(B, w.shape[0], L)


def conv_transpose2x2(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
                      skip: Optional[torch.Tensor] = None) -> torch.Tensor:
    if _grad_on(x, w, bias, skip):
        from . import autograd as ag
        return ag.conv_transpose2x2(x, w, bias, skip)
    x, w = _ck(x, "x"), _ck(w, "w")
    N, Cin, H, W = x.shape
    Cout = w.shape[1]
    bias = None if bias is None else _ck(bias, "bias")
    skip = None if skip is None else _ck(skip, "skip")
    y = torch.empty((N, Cout, 2 * H, 2 * W), device=x.device, dtype=torch.float32)
    _lib.call("cwfa_convT2x2_f32", x.data_ptr(), w.data_ptr(), _p(bias), _p(skip), y.data_ptr(), N, Cin, H, W, Cout, _stream())
    return y


def depth_stencil3d(x: torch.Tensor, w1, b1, slope, w2, b2) -> torch.Tensor:
    """Conv3d(1->Cm,3,p1) + PReLU + Conv3d(Cm->1,3,p1) over (H,W,depth) of a (B,ch,H,W) tensor."""
    if _grad_on(x, w1, b1, slope, w2, b2):
        from . import autograd as ag
        return ag.depth_stencil3d(x, w1, b1, slope, w2, b2)
    x = _ck(x, "x")
    B, ch, H, W = x.shape
    w1, b1, w2, b2, slope = (_ck(t) for t in (w1, b1, w2, b2, slope))
    Cm = w1.shape[0]
    y = torch.empty_like(x)
    _lib.call("cwfa_depth_stencil3d_f32", x.data_ptr(), w1.data_ptr(), b1.data_ptr(), slope.data_ptr(), w2.data_ptr(),
              b2.data_ptr(), y.data_ptr(), B, ch, H, W, Cm, _stream())
    return y


def channel_stats(x: torch.Tensor) -> torch.Tensor:
    """Per-channel (sum, sumsq) over (N,H,W) -> tensor (2,C)."""
    x = _ck(x, "x")
    N, C = x.shape[0], x.shape[1]
    P = x[0, 0].numel()
    stats = torch.empty(2 * C, device=x.device, dtype=torch.float32)
    ws = torch.empty(2 * C * _lib.load().cwfa_stats_workspace_blocks(), device=x.device, dtype=torch.float32)
    _lib.call("cwfa_channel_stats_f32", x.data_ptr(), stats.data_ptr(), ws.data_ptr(), N, C, P, _stream())
    return stats.view(2, C)


def update_running_stats(stats: torch.Tensor, count: float, running_mean, running_var, momentum: Optional[float],
                         num_batches_tracked=None) -> None:
    """nn.BatchNorm2d's training-mode side effect (what the reference's LRNN, left in ``.train()`` mode, does on EVERY forward --
    CWFA.py:531-532, unet.py:100-107 -- and therefore what ends up in its checkpoints):
    ``running = (1 - m) running + m batch`` with the UNBIASED batch variance; ``momentum=None`` = cumulative average.
    ``stats`` = per-channel (sum, sum of squares) from the statistics kernel; C-element parameter algebra on the host side."""
    if running_mean is None or running_var is None:
        return
    C = running_mean.numel()
    s, q = stats[:C].double(), stats[C:].double()
    mean = s / count
    var_unbiased = (q - s * mean) / max(count - 1.0, 1.0)
    with torch.no_grad():
        if num_batches_tracked is not None:
            num_batches_tracked += 1
        m = momentum if momentum is not None else 1.0 / float(num_batches_tracked if num_batches_tracked is not None else 1)
        running_mean.mul_(1.0 - m).add_(mean.to(running_mean.dtype), alpha=m)
        running_var.mul_(1.0 - m).add_(var_unbiased.to(running_var.dtype), alpha=m)


def batchnorm(x: torch.Tensor, gamma, beta, running_mean=None, running_var=None, *, batch_stats: bool,
              eps: float = 1e-5, momentum: Optional[float] = None, update_running: bool = False,
              num_batches_tracked=None) -> torch.Tensor:
    """nn.BatchNorm2d forward: batch statistics (training-mode normalisation, what the reference's
    LRNN runs at inference, CWFA.py:531-532) or running statistics (eval mode).  ``update_running``: also apply the
    training-mode running-statistics update (``update_running_stats``)."""
    if _grad_on(x, gamma, beta):
        from . import autograd as ag
        return ag.batchnorm(x, gamma, beta, running_mean, running_var, batch_stats=batch_stats, eps=eps,
                            momentum=momentum, update_running=update_running, num_batches_tracked=num_batches_tracked)
    x = _ck(x, "x")
    N, C = x.shape[0], x.shape[1]
    P = x[0, 0].numel()
    scale = torch.empty(C, device=x.device, dtype=torch.float32)
    shift = torch.empty(C, device=x.device, dtype=torch.float32)
    if batch_stats:
        stats = channel_stats(x).reshape(-1)
        count = float(N * P)
        if update_running:
            update_running_stats(stats, count, running_mean, running_var, momentum, num_batches_tracked)
    else:
        rm, rv = _ck(running_mean), _ck(running_var)
        stats = torch.cat([rm, rv + rm * rm]).contiguous()     # as (sum, sumsq) with count 1
        count = 1.0
    _lib.call("cwfa_bn_finalize_f32", stats.data_ptr(), _ck(gamma).data_ptr(), _ck(beta).data_ptr(), scale.data_ptr(),
              shift.data_ptr(), C, count, float(eps), _stream())
    y = torch.empty_like(x)
    _lib.call("cwfa_scale_shift_f32", x.data_ptr(), scale.data_ptr(), shift.data_ptr(), y.data_ptr(), N, C, P, _stream())
    return y


def scale_shift(x: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor) -> torch.Tensor:
    """y[b,c,...] = x[b,c,...] * scale[c] + shift[c]  (ActNorm / global affine of AllInOneBlock)."""
    if _grad_on(x, scale, shift):
        from . import autograd as ag
        return ag.scale_shift(x, scale, shift)
    x = _ck(x, "x")
    N, C = x.shape[0], x.shape[1]
    P = x[0, 0].numel()
    y = torch.empty_like(x)
    _lib.call("cwfa_scale_shift_f32", x.data_ptr(), _ck(scale, "scale").data_ptr(), _ck(shift, "shift").data_ptr(), y.data_ptr(), N, C, P, _stream())
    return y


def maxpool2(x: torch.Tensor) -> torch.Tensor:
    if _grad_on(x):
        from . import autograd as ag
        return ag.maxpool2(x)
    x = _ck(x, "x")
    N, C, H, W = x.shape
    y = torch.empty((N, C, H // 2, W // 2), device=x.device, dtype=torch.float32)
    _lib.call("cwfa_maxpool2_f32", x.data_ptr(), y.data_ptr(), N, C, H, W, _stream())
    return y


def layernorm_chw(x: torch.Tensor, gamma, beta, eps: float = 1e-5) -> torch.Tensor:
    if _grad_on(x, gamma, beta):
        from . import autograd as ag
        return ag.layernorm_chw(x, gamma, beta, eps)
    x = _ck(x, "x")
    N = x.shape[0]
    n = x[0].numel()
    y = torch.empty_like(x)
    ws = torch.empty(2 * N * _lib.load().cwfa_layernorm_workspace_blocks(), device=x.device, dtype=torch.float32)
    _lib.call("cwfa_layernorm_chw_f32", x.data_ptr(), _ck(gamma).data_ptr(), _ck(beta).data_ptr(), y.data_ptr(),
              ws.data_ptr(), N, n, float(eps), _stream())
    return y


def gate_add_(x: torch.Tensor, m: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """x += m * 2 * (g - 0.5) in place (networks.py:554)."""
    m, g = _ck(m, "m"), _ck(g, "g")
    if not (x.is_cuda and x.is_contiguous() and x.dtype == torch.float32):
        raise RuntimeError("gate_add_: x must be a contiguous fp32 CUDA tensor")
    _lib.call("cwfa_gate_add_f32", x.data_ptr(), m.data_ptr(), g.data_ptr(), x.numel(), _stream())
    return x


def sum_squares(x: torch.Tensor) -> torch.Tensor:
    """Per-sample sum of squares over all non-batch axes -> (B,) (the ||z||^2 of CWFA.py:183)."""
    if _grad_on(x):
        from . import autograd as ag
        return ag.sum_squares(x)
    x = _ck(x, "x")
    B = x.shape[0]
    n = x[0].numel()
    # samples play the role of channels: view as (N=1, C=B*K, P=n/K).  K pseudo-channels per sample keep the reduction's grid
    # (32 blocks per channel) wide enough for a whole GPU at small batch (K = 1 took 340 us for one 512x512x96 sample, K = 64
    # runs at the HBM rate); the K partial sums are added in a fixed order.
    K = 64 if (B <= 16 and n % 64 == 0 and n >= (1 << 16)) else 1
    C = B * K
    stats = torch.empty(2 * C, device=x.device, dtype=torch.float32)
    ws = torch.empty(2 * C * _lib.load().cwfa_stats_workspace_blocks(), device=x.device, dtype=torch.float32)
    _lib.call("cwfa_channel_stats_f32", x.data_ptr(), stats.data_ptr(), ws.data_ptr(), 1, C, n // K, _stream())
    return stats[C:] if K == 1 else stats[C:].view(B, K).sum(1)


def cast_f16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 -> fp16 copy of a finished volume into ``out`` (device); used for half-size device->host transfers."""
    x = _ck(x, "x")
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.float16)
    if not (out.is_cuda and out.is_contiguous() and out.dtype == torch.float16 and out.numel() == x.numel()):
        raise ValueError("cast_f16: out must be a contiguous fp16 CUDA tensor of the same size")
    _lib.call("cwfa_cast_f32_f16", x.data_ptr(), out.data_ptr(), x.numel(), _stream())
    return out


def batch_mean(x: torch.Tensor) -> torch.Tensor:
    """Mean over the batch axis, keepdim: (K,...) -> (1,...) (the multi-sample average of CWFA.py:913-914), as K - 1 axpby passes."""
    x = _ck(x, "x")
    K = x.shape[0]
    n = x[0].numel()
    bufs = [torch.empty((1,) + tuple(x.shape[1:]), device=x.device, dtype=torch.float32) for _ in range(2 if K > 1 else 1)]
    st = _stream()
    _lib.call("cwfa_axpby_f32", x.data_ptr(), None, bufs[0].data_ptr(), 1.0 / K, 0.0, n, st)
    for k in range(1, K):                                     # ping-pong: the kernel's operands are __restrict__
        _lib.call("cwfa_axpby_f32", x.data_ptr() + 4 * k * n, bufs[(k - 1) & 1].data_ptr(), bufs[k & 1].data_ptr(), 1.0 / K, 1.0, n, st)
    return bufs[(K - 1) & 1]


def attention_gate_(x: torch.Tensor, m: torch.Tensor, v: torch.Tensor, att) -> torch.Tensor:
    """x += m * 2 * (GlobalAttention(v) - 0.5) in one kernel (networks.py:244-262, :554).  ``att`` is the GlobalAttention
    module (parameter holder: att.m[0] Conv1d k=3, att.m[2] Conv1d k=1)."""
    m, v = _ck(m, "m"), _ck(v, "v")
    if not (x.is_cuda and x.is_contiguous() and x.dtype == torch.float32):
        raise RuntimeError("attention_gate_: x must be a contiguous fp32 CUDA tensor")
    B, C = v.shape[0], v.shape[1]
    L = v[0, 0].numel()
    w1, b1, w2, b2 = (_ck(t.detach()) for t in (att.m[0].weight, att.m[0].bias, att.m[2].weight, att.m[2].bias))
    _lib.call("cwfa_attention_gate_f32", x.data_ptr(), m.data_ptr(), v.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
              b2.data_ptr(), B, C, L, _stream())
    return x

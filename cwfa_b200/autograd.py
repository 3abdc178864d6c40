"""Differentiable versions of the fp32 module-path ops: ``torch.autograd.Function`` is used only as the
tape (which op ran, what it saved); every forward AND every adjoint below is one of this repo's CUDA
kernels called through the C ABI (csrc/backward.cu, elementwise.cu, conv_f32.cu).

The reference trains a flow level with stock autograd (CWFA.py:928-1015): an inverse pass with gradients for
the MSE term, a forward pass for the NLL term, ``backward()`` and a Lion step.  ``cwfa_b200.ops`` routes here
whenever gradients are enabled and an input requires them, so the FrEIA-style modules
(``cwfa_b200.modules`` / ``networks``) train unchanged.  There is no CPU fallback.

Covered (everything a flow level and its conditioning net execute): conv2d 1x1/3x3 (+bias, ELU, residual-add,
PReLU), the affine coupling with log-det, channel / row / column permutations, the depth-wise Haar transform
(split / merge forms included), per-sample sum of squares, MSE, the conditioning net's depth stencil.
The LRNN's U-Net is covered too (BatchNorm2d, 2x2 max-pool, ConvTranspose2d(k=2,s=2) + skip add as a 1x1 convolution followed by
a pixel shuffle) and so is its mean-volume branch (7x7 conv, LayerNorm([C,H,W]), GELU + skip, attention gate).  Ops without an
adjoint (e.g. the fused depth-wise attention kernel, 2-D Haar) stay non-differentiable and ``ops`` says so once.  ``set_training_precision('bf16'|'fp16')`` moves every convolution of the tape
(forward, data gradient, weight gradient) onto the tcgen05 kernels.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib, ops

_F = torch.autograd.Function


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32(t: torch.Tensor) -> torch.Tensor:
    return ops._ck(t.detach())


def _zeros_like_if_none(g, ref):
    return torch.zeros_like(ref) if g is None else ops._ck(g)


# ---------------------------------------------------------------------------------------------
# raw adjoint launches
# ---------------------------------------------------------------------------------------------
def conv2d_wgrad(x: torch.Tensor, dy: torch.Tensor, KH: int, KW: int) -> torch.Tensor:
    """dW (Cout,Cin,KH,KW) of a stride-1 'same' convolution from its input and output cotangent."""
    N, Cin, H, W = x.shape
    Cout = dy.shape[1]
    lib = _lib.load()
    nws = lib.cwfa_conv2d_wgrad_workspace_floats(N, Cin, H, W, Cout, KH, KW)
    ws = torch.empty(nws, device=x.device, dtype=torch.float32)
    dw = torch.empty((Cout, Cin, KH, KW), device=x.device, dtype=torch.float32)
    _lib.call("cwfa_conv2d_wgrad_f32", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), ws.data_ptr(), N, Cin, H, W, Cout, KH, KW, 0, _stream())
    return dw


def conv2d_dgrad(dy: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """dx = conv_same(dy, flip(w)^T): the forward direct-convolution kernel on re-laid weights."""
    Cout, Cin, KH, KW = w.shape
    wt = torch.empty((Cin, Cout, KH, KW), device=w.device, dtype=torch.float32)
    _lib.call("cwfa_conv2d_dgrad_weights_f32", w.data_ptr(), wt.data_ptr(), Cout, Cin, KH, KW, _stream())
    return ops.conv2d(dy, wt, None)


def channel_sum(x: torch.Tensor) -> torch.Tensor:
    """Per-channel sum over (N,H,W) of a (N,C,...) tensor (bias gradients)."""
    return ops.channel_stats(x)[0]


def axpby(a: torch.Tensor, b: Optional[torch.Tensor], alpha: float, beta: float) -> torch.Tensor:
    out = torch.empty_like(a)
    _lib.call("cwfa_axpby_f32", a.data_ptr(), ops._p(b), out.data_ptr(), float(alpha), float(beta), a.numel(), _stream())
    return out


def scale_per_sample(x: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """y[b] = s[b] * x[b] with s a (B,) device tensor (no host sync)."""
    B = x.shape[0]
    n = x[0].numel()
    y = torch.empty_like(x)
    s = ops._ck(s).reshape(B)
    _lib.call("cwfa_scale_shift_f32", x.data_ptr(), s.data_ptr(), torch.zeros_like(s).data_ptr(), y.data_ptr(), 1, B, n, _stream())
    return y


# ---------------------------------------------------------------------------------------------
# conv2d (+bias, +residual before the activation, ELU fused; PReLU as its own op)
# ---------------------------------------------------------------------------------------------
class _Conv2d(_F):
    @staticmethod
    def forward(ctx, x, w, bias, res, act, res_mode):
        xx, ww = _f32(x), _f32(w)
        bb = None if bias is None else _f32(bias)
        rr = None if res is None else _f32(res)
        y = ops.conv2d(xx, ww, bb, act=act, res=rr, res_mode=res_mode)
        ctx.act = act
        ctx.has_res = rr is not None and res_mode == 1
        ctx.save_for_backward(xx, ww, y if act != ops.ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dv = ops._ck(dy)
        if ctx.act != ops.ACT_NONE:                     # ELU / ReLU / sigmoid: adjoint from the layer output
            g = torch.empty_like(dv)
            _lib.call("cwfa_act_bwd_f32", dv.data_ptr(), y.data_ptr(), g.data_ptr(), dv.numel(), int(ctx.act), _stream())
            dv = g
        need_x, need_w, need_b, need_r = ctx.needs_input_grad[:4]
        dx = conv2d_dgrad(dv, w) if need_x else None
        dw = conv2d_wgrad(x, dv, w.shape[2], w.shape[3]) if need_w else None
        db = channel_sum(dv) if need_b else None
        dr = dv if (need_r and ctx.has_res) else None
        return dx, dw, db, dr, None, None


class _PReLU(_F):
    @staticmethod
    def forward(ctx, v, slope):
        vv, ss = _f32(v), _f32(slope)
        if ss.numel() != 1:
            raise NotImplementedError("cwfa_b200: PReLU adjoint implements the single shared slope CWFA uses (networks.py:209)")
        y = torch.empty_like(vv)
        _lib.call("cwfa_prelu_f32", vv.data_ptr(), ss.data_ptr(), y.data_ptr(), vv.numel(), _stream())
        ctx.save_for_backward(vv, ss)
        return y

    @staticmethod
    def backward(ctx, dy):
        v, slope = ctx.saved_tensors
        dy = ops._ck(dy)
        dv = torch.empty_like(v)
        ds = torch.empty(1, device=v.device, dtype=torch.float32)
        ws = torch.empty(_lib.load().cwfa_reduce_workspace_blocks(), device=v.device, dtype=torch.float32)
        _lib.call("cwfa_prelu_bwd_f32", dy.data_ptr(), v.data_ptr(), slope.data_ptr(), dv.data_ptr(), ds.data_ptr(), ws.data_ptr(),
                  v.numel(), _stream())
        return dv, ds.reshape(slope.shape)


def prelu(v, slope):
    return _PReLU.apply(v, slope)


_FROM_OUTPUT = (ops.ACT_NONE, ops.ACT_ELU, ops.ACT_RELU, ops.ACT_SIGMOID)


def conv2d_supported(w, act, res, res_mode) -> bool:
    KH, KW = w.shape[2], w.shape[3]
    if KH != KW or KH not in (1, 3, 7):
        return False
    if act == ops.ACT_GELU:                             # ConvNeXt tail: gelu(conv) + skip (networks.py:491-503)
        return res is None or res_mode == 2
    return act in _FROM_OUTPUT + (ops.ACT_PRELU,) and (res is None or res_mode == 1)


# ---- tensor-core variant: forward and data gradient on the tcgen05 implicit-GEMM kernel (bf16/fp16 operands, fp32
# accumulation and fp32 NCHW results), weight gradient in fp32.  BASELINE.json configs[3] names bf16 for the training step;
# the reference's own GPU path trains under fp16 autocast (CWFA.py:845,996).
_PRECISION = "fp32"


def set_training_precision(kind: str) -> str:
    """'fp32' (reference precision, direct CUDA-core convolutions) or 'bf16' / 'fp16' (tensor-core convolutions in the
    differentiable path).  Returns the previous setting."""
    global _PRECISION
    if kind not in ("fp32", "bf16", "fp16"):
        raise ValueError(f"unknown training precision {kind!r}")
    prev, _PRECISION = _PRECISION, kind
    return prev


def training_precision() -> str:
    return _PRECISION


def conv2d_wgrad_tc(x8, dy8, Cin: int, Cout: int, K: int) -> torch.Tensor:
    """dW (Cout,Cin,K,K) fp32 from C8 half-precision x and dy on the tensor cores (csrc/wgrad_tc.cu)."""
    lib = _lib.load()
    N, H, W = x8.N, x8.H, x8.W
    nws = lib.cwfa_wgrad_tc_workspace_floats(N, H, W, Cin, x8.Cp, Cout, dy8.Cp, K)
    ws = torch.empty(nws, device=x8.data.device, dtype=torch.float32)
    dw = torch.empty((Cout, Cin, K, K), device=x8.data.device, dtype=torch.float32)
    _lib.call("cwfa_wgrad_tc", x8.data.data_ptr(), dy8.data.data_ptr(), dw.data_ptr(), ws.data_ptr(), N, H, W, Cin, x8.Cp, Cout, dy8.Cp,
              K, K, x8.is_bf16, _stream())
    return dw


def _packed_for(w, bias, kind: str, transposed_for_dgrad: bool = False):
    """Packed tensor-core weights of a parameter, cached ON the parameter object: re-packed only when the parameter changed
    (in-place version / storage) or after an optimiser step (``packed.weights_changed``, called by ``Lion.step``) -- a training
    step otherwise spends ~6 % of its time re-packing the same weights for the forward conv and for the data gradient."""
    from . import packed, tc
    b_key = None if bias is None else (bias._version, bias.data_ptr())
    key = (kind, packed._EPOCH, w._version, w.data_ptr(), b_key, transposed_for_dgrad)
    cache = getattr(w, "_cwfa_pack", None)
    if cache is None:
        cache = {}
        try:
            w._cwfa_pack = cache
        except AttributeError:
            pass
    slot = "dgrad" if transposed_for_dgrad else "fwd"
    hit = cache.get(slot)
    if hit is not None and hit[0] == key:
        return hit[1]
    ww = _f32(w)
    if transposed_for_dgrad:
        Cout, Cin, KH, KW = ww.shape
        wt = torch.empty((Cin, Cout, KH, KW), device=ww.device, dtype=torch.float32)
        _lib.call("cwfa_conv2d_dgrad_weights_f32", ww.data_ptr(), wt.data_ptr(), Cout, Cin, KH, KW, _stream())
        pc = tc.PackedConv(wt, None, kind)
    else:
        pc = tc.PackedConv(ww, None if bias is None else bias.detach(), kind)
    cache[slot] = (key, pc)
    return pc


class _Conv2dTC(_F):
    @staticmethod
    def forward(ctx, x, w, bias, res, act, res_mode, kind):
        from . import tc
        xx, ww = _f32(x), _f32(w)
        rr = None if res is None else _f32(res)
        pc = _packed_for(w, bias, kind)
        ctx.w_ref = w
        x8 = tc.to_c8(xx, kind)
        y = tc.conv_tc(x8, pc, act=act, res=rr, res_mode=res_mode, out_nchw=True)
        ctx.act, ctx.kind = act, kind
        ctx.has_res = rr is not None and res_mode == 1
        ctx.wgrad_tc = ww.shape[2] in (1, 3)                             # 7x7: fp32 weight gradient (csrc/backward.cu)
        ctx.Cin = xx.shape[1]
        ctx.save_for_backward(None if ctx.wgrad_tc else xx, ww, y if act == ops.ACT_ELU else None, x8.data if ctx.wgrad_tc else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import tc
        x, w, y, x8d = ctx.saved_tensors
        need_x, need_w, need_b, need_r = ctx.needs_input_grad[:4]
        Cout, Cin, KH, KW = w.shape
        elu = ctx.act == ops.ACT_ELU
        need_f32 = (need_r and ctx.has_res) or (need_w and not ctx.wgrad_tc)       # residual cotangent / fp32 weight gradient
        # ONE pass over dy: ELU adjoint, C8 conversion for the MMAs, bias gradient (csrc/backward.cu: dy_prep_kernel)
        dv8, dv, db = tc.dy_prep(dy, y if elu else None, ctx.kind, want_f32=need_f32 and elu, want_bias=need_b)
        if need_f32 and not elu:
            dv = ops._ck(dy)
        dx = None
        if need_x:
            dx = tc.conv_tc(dv8, _packed_for(ctx.w_ref, None, ctx.kind, transposed_for_dgrad=True), out_nchw=True)
        dw = None
        if need_w:
            dw = conv2d_wgrad_tc(tc.C8(x8d, Cin, ctx.kind), dv8, Cin, Cout, KH) if ctx.wgrad_tc else conv2d_wgrad(x, dv, KH, KW)
        dr = dv if (need_r and ctx.has_res) else None
        return dx, dw, db, dr, None, None, None


class _SubnetTC(_F):
    """A whole coupling sub-network (1x1 in -> 3 x [3x3 + ELU, 1x1 + skip + ELU] -> 3x3 out, networks.py:624-638,659-671) as ONE
    autograd node whose activations AND cotangents stay in the C8 half layout: every convolution reads and writes C8 (ELU and
    the skip add in its epilogue), the adjoint walks back with ``elu_bwd_c8`` (+ bias sums) -> data-gradient conv (skip add in
    the epilogue) -> tensor-core weight gradient from the saved C8 operands.  The per-convolution node chain converted an fp32
    NCHW tensor to C8 before every convolution and after every adjoint (12 % of a training step)."""

    @staticmethod
    def forward(ctx, inp, kind, *params):
        from . import tc
        (w_in, b_in), blocks, (w_out, b_out) = (params[0], params[1]), [params[2 + 4 * k: 6 + 4 * k] for k in range(3)], (params[14], params[15])
        x8 = tc.to_c8(_f32(inp), kind)
        b = tc.conv_tc(x8, _packed_for(w_in, b_in, kind))
        saved = [x8.data, b.data]
        for (w3, b3, w1, b1) in blocks:
            t = tc.conv_tc(b, _packed_for(w3, b3, kind), act=ops.ACT_ELU)
            b = tc.conv_tc(t, _packed_for(w1, b1, kind), act=ops.ACT_ELU, res=b, res_mode=1)
            saved += [t.data, b.data]
        y = tc.conv_tc(b, _packed_for(w_out, b_out, kind), out_nchw=True)
        ctx.save_for_backward(*saved)
        ctx.params, ctx.kind, ctx.Cin, ctx.n = params, kind, inp.shape[1], w_in.shape[0]
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import tc
        kind, n, Cin, params = ctx.kind, ctx.n, ctx.Cin, ctx.params
        sv = ctx.saved_tensors
        C8 = lambda d, c: tc.C8(d, c, kind)
        x8, bs, ts = C8(sv[0], Cin), [C8(sv[1 + 2 * k], n) for k in range(4)], [C8(sv[2 + 2 * k], n) for k in range(3)]
        w_out = params[14]
        grads = [None] * 16
        dy8, _, grads[15] = tc.dy_prep(dy, None, kind, want_bias=True)
        grads[14] = conv2d_wgrad_tc(bs[3], dy8, n, w_out.shape[0], w_out.shape[2])
        d = tc.conv_tc(dy8, _packed_for(w_out, None, kind, transposed_for_dgrad=True))            # cotangent of b_3 (post-ELU)
        for k in (2, 1, 0):
            w3, b3, w1, b1 = params[2 + 4 * k: 6 + 4 * k]
            g, st = tc.elu_bwd_c8(d, bs[k + 1])                                                   # through ELU(conv1x1(t) + b_k)
            grads[5 + 4 * k] = st[:n].clone()
            grads[4 + 4 * k] = conv2d_wgrad_tc(ts[k], g, n, n, 1)
            dt = tc.conv_tc(g, _packed_for(w1, None, kind, transposed_for_dgrad=True))
            gt, st = tc.elu_bwd_c8(dt, ts[k])                                                     # through ELU(conv3x3(b_k))
            grads[3 + 4 * k] = st[:n].clone()
            grads[2 + 4 * k] = conv2d_wgrad_tc(bs[k], gt, n, n, 3)
            d = tc.conv_tc(gt, _packed_for(w3, None, kind, transposed_for_dgrad=True), res=g, res_mode=1)   # + the skip branch
        w_in = params[0]
        grads[1] = tc.channel_sums_c8(d)[:n]
        grads[0] = conv2d_wgrad_tc(x8, d, Cin, n, w_in.shape[2])
        dx = tc.conv_tc(d, _packed_for(w_in, None, kind, transposed_for_dgrad=True), out_nchw=True) if ctx.needs_input_grad[0] else None
        return (dx, None) + tuple(grads)


# CWFA_SUBNET_NODE=0 keeps one autograd node per convolution (A/B measurements)
_SUBNET_NODE = __import__("os").environ.get("CWFA_SUBNET_NODE", "1") == "1"


def subnet_tc_supported(inp, convs) -> bool:
    """The C8-native sub-network node needs a half training precision, CUDA tensors, 64 internal channels' worth of odd
    square kernels with biases -- the reference's sub-network."""
    if not _SUBNET_NODE or _PRECISION == "fp32" or not torch.is_grad_enabled() or not inp.is_cuda or inp.dim() != 4:
        return False
    return all(c.bias is not None and c.weight.shape[2] == c.weight.shape[3] and c.weight.shape[2] in (1, 3) for c in convs)


def subnet_tc(inp, conv_in, blocks, conv_out):
    """``conv_out(trunk(conv_in(inp)))`` of wavelet_flow_subnetwork2D through the single C8-native node."""
    params = [conv_in.weight, conv_in.bias]
    for c3, c1 in blocks:
        params += [c3.weight, c3.bias, c1.weight, c1.bias]
    params += [conv_out.weight, conv_out.bias]
    return _SubnetTC.apply(inp, _PRECISION, *params)


class _GeluAdd(_F):
    """y = gelu(v) + r (exact erf GELU + the ConvNeXt skip, networks.py:492,503)."""

    @staticmethod
    def forward(ctx, v, r):
        vv = _f32(v)
        rr = None if r is None else _f32(r)
        ctx.save_for_backward(vv)
        y = torch.empty_like(vv)
        _lib.call("cwfa_gelu_add_f32", vv.data_ptr(), ops._p(rr), None, y.data_ptr(), vv.numel(), _stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        (v,) = ctx.saved_tensors
        dy = ops._ck(dy)
        dv = torch.empty_like(v)
        _lib.call("cwfa_gelu_add_f32", v.data_ptr(), None, dy.data_ptr(), dv.data_ptr(), v.numel(), _stream())
        return dv, (dy if ctx.needs_input_grad[1] else None)


def conv2d(x, w, bias=None, *, act=ops.ACT_NONE, slope=None, res=None, res_mode=0):
    rm = res_mode if res is not None else 0
    if act == ops.ACT_GELU:
        lin = _Conv2d.apply(x, w, bias, None, ops.ACT_NONE, 0) if _PRECISION == "fp32" else \
            _Conv2dTC.apply(x, w, bias, None, ops.ACT_NONE, 0, _PRECISION)
        return _GeluAdd.apply(lin, res)
    a = ops.ACT_NONE if act == ops.ACT_PRELU else act
    if _PRECISION == "fp32" or a in (ops.ACT_RELU, ops.ACT_SIGMOID):     # the tiny attention convs stay on the fp32 kernels
        y = _Conv2d.apply(x, w, bias, res, a, rm)
    else:
        y = _Conv2dTC.apply(x, w, bias, res, a, rm, _PRECISION)
    return prelu(y, slope) if act == ops.ACT_PRELU else y


def conv1d_flat(x, w, bias, act):
    """Conv1d over the flattened H*W axis (GlobalAttention, networks.py:250-262) as a convolution on a one-row image; a k = 3
    kernel is embedded in the middle row of a 3x3 kernel (rows above / below the image are zero padding)."""
    B, C, L = x.shape
    k = w.shape[-1]
    w4 = w.reshape(w.shape[0], w.shape[1], 1, k)
    if k == 3:
        w4 = torch.nn.functional.pad(w4, (0, 0, 1, 1))
    elif k != 1:
        raise NotImplementedError("conv1d_flat adjoint: kernel sizes 1 and 3")
    return conv2d(x.reshape(B, C, 1, L), w4, bias, act=act).reshape(B, w.shape[0], L)


class _LayerNormCHW(_F):
    """LayerNorm([C,H,W]) with element-wise affine (networks.py:490)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        xx, gg = _f32(x), _f32(gamma)
        ctx.eps = float(eps)
        ctx.save_for_backward(xx, gg)
        return ops.layernorm_chw(xx, gg, _f32(beta), eps)

    @staticmethod
    def backward(ctx, dy):
        x, gamma = ctx.saved_tensors
        dy = ops._ck(dy)
        B = x.shape[0]
        n = x[0].numel()
        dev = x.device
        s, q = ops.channel_stats(x.reshape(1, B, n)).double()             # per-sample sum, sum of squares
        mean = s / n
        rstd = 1.0 / torch.sqrt((q / n - mean * mean).clamp_min(0.0) + ctx.eps)
        lib = _lib.load()
        out = torch.empty(2 * B, device=dev, dtype=torch.float32)
        ws = torch.empty(2 * B * lib.cwfa_channel_dot_workspace_blocks(), device=dev, dtype=torch.float32)
        _lib.call("cwfa_ln_bwd_stats_f32", x.data_ptr(), dy.data_ptr(), gamma.data_ptr(), out.data_ptr(), ws.data_ptr(), B, n, _stream())
        sg, sgx = out[:B].double(), out[B:].double()
        coef = torch.stack([mean, rstd, sg / n, rstd * (sgx - mean * sg) / n], 1).float().contiguous()
        need_x, need_g, need_b = ctx.needs_input_grad[:3]
        dx = torch.empty_like(x) if need_x else None
        dg = torch.empty_like(gamma) if need_g else None
        db = torch.empty_like(gamma) if need_b else None
        _lib.call("cwfa_ln_bwd_apply_f32", x.data_ptr(), dy.data_ptr(), gamma.data_ptr(), coef.data_ptr(), ops._p(dx), ops._p(dg), ops._p(db),
                  B, n, _stream())
        return dx, dg, db, None


def layernorm_chw(x, gamma, beta, eps):
    return _LayerNormCHW.apply(x, gamma, beta, eps)


class _GateAdd(_F):
    """y = x + m * 2 * (g - 0.5)  (networks.py:554)."""

    @staticmethod
    def forward(ctx, x, m, g):
        xx, mm, gg = _f32(x), _f32(m), _f32(g)
        ctx.save_for_backward(mm, gg)
        y = torch.empty_like(xx)
        _lib.call("cwfa_gate_f32", xx.data_ptr(), mm.data_ptr(), gg.data_ptr(), None, y.data_ptr(), None, xx.numel(), _stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        m, g = ctx.saved_tensors
        dy = ops._ck(dy)
        need_x, need_m, need_g = ctx.needs_input_grad
        dm = torch.empty_like(m) if need_m else None
        dg = torch.empty_like(g) if need_g else None
        if need_m or need_g:
            _lib.call("cwfa_gate_f32", None, m.data_ptr(), g.data_ptr(), dy.data_ptr(), ops._p(dm), ops._p(dg), m.numel(), _stream())
        return (dy if need_x else None), dm, dg


def gate_add(x, m, g):
    return _GateAdd.apply(x, m, g)


# ---------------------------------------------------------------------------------------------
# affine coupling + log-det
# ---------------------------------------------------------------------------------------------
class _Affine(_F):
    """inputs: x (B,ch,..) or None; a = packed [s_raw | t] (B,2ch,..) when t_src is None, else s_raw (B,ch,..) with the
    shift read from t_src (B,ch,..).  outputs: y, logdet (B,)."""

    @staticmethod
    def forward(ctx, x, a, t_src, inverse, clamp, t_scale, k_atan, raw):
        aa = _f32(a)
        xx = None if x is None else _f32(x)
        packed = t_src is None
        if packed:
            ch = aa.shape[1] // 2
            a_s, a_t = aa[:, :ch], aa[:, ch:]
        else:
            a_s, a_t = aa, ops._prep_inner(t_src.detach())[0]
            ch = aa.shape[1]
        y, logdet = ops.affine(xx, a_s, a_t, inverse=inverse, clamp=clamp, t_scale=t_scale, k_atan=k_atan, s_is_final=raw)
        ctx.cfg = (bool(inverse), float(clamp), float(t_scale), float(k_atan), int(raw), packed, ch)     # raw: 0 ATAN, 1 final s, 2 TANH
        ctx.save_for_backward(xx, aa, None if packed else a_t)
        return y, logdet

    @staticmethod
    def backward(ctx, dy, g_logdet):
        inverse, clamp, t_scale, k_atan, raw, packed, ch = ctx.cfg
        x, a, t_keep = ctx.saved_tensors
        B = a.shape[0]
        P = a[0, 0].numel()
        n = ch * P
        dev = a.device
        if packed:
            a_s_ptr, a_t_ptr, ld_s, ld_t = a.data_ptr(), a.data_ptr() + 4 * n, 2 * n, 2 * n
        else:
            a_s_ptr, a_t_ptr, ld_s = a.data_ptr(), t_keep.data_ptr(), n
            ld_t = t_keep.stride(0) if B > 1 else n
        dy = torch.zeros((B, ch) + tuple(a.shape[2:]), device=dev, dtype=torch.float32) if dy is None else ops._ck(dy)
        gj = None if g_logdet is None else ops._ck(g_logdet)
        need_x, need_a, need_t = ctx.needs_input_grad[:3]
        dx = torch.empty_like(dy) if (need_x and x is not None) else None
        da = torch.empty_like(a) if need_a else None
        dt = None
        if packed:
            da_s_ptr = None if da is None else da.data_ptr()
            da_t_ptr = None if da is None else da.data_ptr() + 4 * n
            ld_ds = ld_dt = 2 * n
        else:
            da_s_ptr, ld_ds = (None if da is None else da.data_ptr()), n
            dt = torch.empty_like(dy) if need_t else None
            da_t_ptr, ld_dt = (None if dt is None else dt.data_ptr()), n
        _lib.call("cwfa_affine_bwd", ops._p(x), a_s_ptr, a_t_ptr, dy.data_ptr(), ops._p(gj), ops._p(dx), da_s_ptr, da_t_ptr,
                  B, ch, P, ld_s, ld_t, ld_ds, ld_dt, clamp, k_atan, t_scale, int(inverse) | (4 if raw == 2 else (2 if raw else 0)), _stream())
        return dx, da, dt, None, None, None, None, None


def affine(x, a_s, a_t, *, inverse, clamp, t_scale, k_atan, s_is_final):
    """Differentiable ``ops.affine``: returns (y, logdet).  Detects the packed [s|t] sub-network output so that its
    gradient is written in place instead of being assembled from two slice gradients."""
    ch = a_s.shape[1]
    base = a_s._base if a_s._base is not None else None
    if (base is not None and a_t._base is base and base.dim() == a_s.dim() and base.is_contiguous() and base.shape[1] == 2 * ch
            and base.dtype == torch.float32 and a_s.data_ptr() == base.data_ptr()
            and a_t.data_ptr() == base.data_ptr() + 4 * ch * a_s[0, 0].numel()):
        return _Affine.apply(x, base, None, inverse, clamp, t_scale, k_atan, s_is_final)
    return _Affine.apply(x, a_s.contiguous(), a_t, inverse, clamp, t_scale, k_atan, s_is_final)


# ---------------------------------------------------------------------------------------------
# permutations, Haar, reductions
# ---------------------------------------------------------------------------------------------
def _inverse_perm(perm: torch.Tensor) -> torch.Tensor:
    cache = getattr(perm, "_cwfa_inv", None)
    key = perm._version
    if cache is None or cache[0] != key:
        inv = torch.empty_like(perm)
        inv[perm.detach()] = torch.arange(perm.numel(), device=perm.device, dtype=perm.dtype)
        cache = (key, inv)
        try:
            perm._cwfa_inv = cache
        except AttributeError:
            pass
    return cache[1]


class _Permute(_F):
    @staticmethod
    def forward(ctx, x, perm, axis):
        ctx.perm, ctx.axis = perm, axis
        return ops.permute(_f32(x), perm, axis)

    @staticmethod
    def backward(ctx, dy):
        return ops.permute(ops._ck(dy), _inverse_perm(ctx.perm), ctx.axis), None, None


def permute(x, perm, axis):
    return _Permute.apply(x, perm, axis)


class _Haar1d(_F):
    """The orthonormal depth-wise Haar transform: the adjoint of each direction is the other direction."""

    @staticmethod
    def forward(ctx, x, inverse):
        ctx.inverse = inverse
        xx = _f32(x)
        return ops.haar1d_inverse(xx) if inverse else ops.haar1d_forward(xx)

    @staticmethod
    def backward(ctx, dy):
        dy = ops._ck(dy)
        return (ops.haar1d_forward(dy) if ctx.inverse else ops.haar1d_inverse(dy)), None


def haar1d(x, inverse: bool):
    return _Haar1d.apply(x, inverse)


class _HaarSplit(_F):
    @staticmethod
    def forward(ctx, x):
        return ops.haar1d_split(_f32(x))

    @staticmethod
    def backward(ctx, dlo, dhi):
        ref = dlo if dlo is not None else dhi
        return ops.haar1d_merge(_zeros_like_if_none(dlo, ref), _zeros_like_if_none(dhi, ref))


class _HaarMerge(_F):
    @staticmethod
    def forward(ctx, lo, hi):
        return ops.haar1d_merge(_f32(lo), _f32(hi))

    @staticmethod
    def backward(ctx, dy):
        return ops.haar1d_split(ops._ck(dy))


def haar1d_split(x):
    return _HaarSplit.apply(x)


def haar1d_merge(lo, hi):
    return _HaarMerge.apply(lo, hi)


class _SumSquares(_F):
    @staticmethod
    def forward(ctx, x):
        xx = _f32(x)
        ctx.save_for_backward(xx)
        return ops.sum_squares(xx).clone()

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return scale_per_sample(x, 2.0 * g)


def sum_squares(x):
    return _SumSquares.apply(x)


class _Mse(_F):
    """F.mse_loss(a, b) (mean over all elements; the regulariser of CWFA.py:951-955)."""

    @staticmethod
    def forward(ctx, a, b):
        aa, bb = _f32(a), _f32(b)
        d = axpby(aa, bb, 1.0, -1.0)
        ctx.save_for_backward(d)
        return ops.sum_squares(d.reshape(1, -1)).reshape(()) / d.numel()

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        s = (2.0 / d.numel()) * g.reshape(1)
        da = scale_per_sample(d.reshape(1, -1), s).reshape(d.shape) if (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]) else None
        ga = da if ctx.needs_input_grad[0] else None
        gb = axpby(da, None, -1.0, 0.0) if ctx.needs_input_grad[1] else None
        return ga, gb


def mse_loss(a, b):
    return _Mse.apply(a, b)


# ---------------------------------------------------------------------------------------------
# conditioning net's depth stencil (networks.py:221-225,239), unfused so the hidden tensors are on the tape
# ---------------------------------------------------------------------------------------------
class _DepthStencil(_F):
    @staticmethod
    def forward(ctx, x, w1, b1, slope, w2, b2):
        xx, ww1, bb1, ss, ww2, bb2 = (_f32(t) for t in (x, w1, b1, slope, w2, b2))
        B, D, H, W = xx.shape
        Cm = ww1.shape[0]
        st = _stream()
        v1 = torch.empty((B, Cm, D, H, W), device=xx.device, dtype=torch.float32)
        _lib.call("cwfa_stencil3d_1toC_f32", xx.data_ptr(), ww1.data_ptr(), bb1.data_ptr(), v1.data_ptr(), B, D, H, W, Cm, 0, st)
        h = torch.empty_like(v1)
        _lib.call("cwfa_prelu_f32", v1.data_ptr(), ss.data_ptr(), h.data_ptr(), v1.numel(), st)
        y = torch.empty_like(xx)
        _lib.call("cwfa_stencil3d_Cto1_f32", h.data_ptr(), ww2.data_ptr(), bb2.data_ptr(), y.data_ptr(), B, D, H, W, Cm, 0, st)
        ctx.save_for_backward(xx, ww1, ss, ww2, v1, h)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w1, slope, w2, v1, h = ctx.saved_tensors
        dy = ops._ck(dy)
        B, D, H, W = x.shape
        Cm = w1.shape[0]
        st = _stream()
        lib = _lib.load()
        dev = x.device
        ws = torch.empty(max(lib.cwfa_stencil3d_wgrad_workspace_floats(Cm), lib.cwfa_reduce_workspace_blocks()), device=dev, dtype=torch.float32)
        dh = torch.empty_like(h)
        _lib.call("cwfa_stencil3d_1toC_f32", dy.data_ptr(), w2.data_ptr(), None, dh.data_ptr(), B, D, H, W, Cm, 1, st)
        dw2 = torch.empty_like(w2)
        _lib.call("cwfa_stencil3d_wgrad_f32", dy.data_ptr(), h.data_ptr(), dw2.data_ptr(), ws.data_ptr(), B, D, H, W, Cm, 1, st)
        db2 = channel_sum(dy.reshape(1, 1, -1)).reshape(1)
        dv1 = dh                                     # in place: dh is not needed afterwards
        dslope = torch.empty(1, device=dev, dtype=torch.float32)
        _lib.call("cwfa_prelu_bwd_f32", dh.data_ptr(), v1.data_ptr(), slope.data_ptr(), dv1.data_ptr(), dslope.data_ptr(), ws.data_ptr(),
                  v1.numel(), st)
        dw1 = torch.empty_like(w1)
        _lib.call("cwfa_stencil3d_wgrad_f32", x.data_ptr(), dv1.data_ptr(), dw1.data_ptr(), ws.data_ptr(), B, D, H, W, Cm, 0, st)
        db1 = channel_sum(dv1.reshape(B, Cm, -1))
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _lib.call("cwfa_stencil3d_Cto1_f32", dv1.data_ptr(), w1.data_ptr(), None, dx.data_ptr(), B, D, H, W, Cm, 1, st)
        return dx, dw1, db1, dslope.reshape(slope.shape), dw2, db2


_BAND = {}


def _band_index(D: int, device):
    """kd[d, d'] = d' - d + 1 clipped to [0, 2] and mask[d, d'] = (|d' - d| <= 1): zero padding in depth falls out of the band."""
    key = (D, str(device))
    if key not in _BAND:
        d = torch.arange(D, device=device)
        off = d[None, :] - d[:, None] + 1
        _BAND[key] = (off.clamp(0, 2).reshape(-1), ((off >= 0) & (off <= 2)).float())
    return _BAND[key]


def depth_stencil3d_banded(x, w1, b1, slope, w2, b2):
    """The depth stencil as two ordinary 3x3 2-D convolutions on the tensor cores (the formulation of the inference engine,
    cwfa_b200/engine.py:_CondNet): channels carry (depth, hidden channel), weights are the depth-banded expansion
    W1[(d,c), d'] = w1[c,:,:,d'-d+1], W2[d, (d',c)] = w2[c,:,:,d'-d+1].  The expansion is a gather of the
    (Cm,27) parameters, so autograd folds the banded weight gradients (wgrad_tc) back onto them."""
    D, Cm = x.shape[1], w1.shape[0]
    kd, mask = _band_index(D, x.device)
    # gather along kd (a pure re-indexing of the (Cm,27) parameters; its adjoint is torch's index_add)
    g1 = w1[:, 0].index_select(3, kd).reshape(Cm, 3, 3, D, D) * mask            # [c, h, w, d, d']
    g2 = w2[0].index_select(3, kd).reshape(Cm, 3, 3, D, D) * mask
    W1 = g1.permute(3, 0, 4, 1, 2).reshape(D * Cm, D, 3, 3)                      # [(d,c), d', h, w]
    W2 = g2.permute(3, 4, 0, 1, 2).reshape(D, D * Cm, 3, 3)                      # [d, (d',c), h, w]
    hid = conv2d(x, W1, b1.repeat(D), act=ops.ACT_PRELU, slope=slope)
    return conv2d(hid, W2, b2.expand(D))


def _dgrad_weights(w: torch.Tensor) -> torch.Tensor:
    """(Cout,Cin,K,K) -> the flipped / transposed (Cin,Cout,K,K) weights whose forward convolution is the data gradient."""
    Cout, Cin, KH, KW = w.shape
    wt = torch.empty((Cin, Cout, KH, KW), device=w.device, dtype=torch.float32)
    _lib.call("cwfa_conv2d_dgrad_weights_f32", w.data_ptr(), wt.data_ptr(), Cout, Cin, KH, KW, _stream())
    return wt


class _StencilBandedTC(_F):
    """The banded depth stencil of ``depth_stencil3d_banded`` as ONE autograd node whose 32 D-channel hidden tensor lives only
    in the C8 half layout: conv (C8 out) -> PReLU on C8 -> conv, and in the adjoint data gradient (C8 out) -> PReLU adjoint on
    C8 (+ bias / slope sums) -> the two tensor-core weight gradients straight from the saved C8 operands.  The generic node
    chain moved the 1.6 GB fp32 NCHW form of that tensor six times per step (conv epilogue, PReLU, conversion, and the same
    three on the way back).  Banded weight gradients are folded back onto the (Cm, 27) parameters here (adjoint of the gather)."""

    @staticmethod
    def forward(ctx, x, w1, b1, slope, w2, b2, kind):
        from . import tc
        D, Cm = x.shape[1], w1.shape[0]
        kd, mask = _band_index(D, x.device)
        g1 = _f32(w1)[:, 0].index_select(3, kd).reshape(Cm, 3, 3, D, D) * mask       # [c, h, w, d, d']
        g2 = _f32(w2)[0].index_select(3, kd).reshape(Cm, 3, 3, D, D) * mask
        W1 = g1.permute(3, 0, 4, 1, 2).reshape(D * Cm, D, 3, 3).contiguous()          # [(d,c), d', h, w]
        W2 = g2.permute(3, 4, 0, 1, 2).reshape(D, D * Cm, 3, 3).contiguous()          # [d, (d',c), h, w]
        x8 = tc.to_c8(_f32(x), kind)
        pre8 = tc.conv_tc(x8, tc.PackedConv(W1, _f32(b1).repeat(D), kind))
        hid8 = tc.prelu_c8(pre8, slope)
        y = tc.conv_tc(hid8, tc.PackedConv(W2, _f32(b2).expand(D).contiguous(), kind), out_nchw=True)
        ctx.save_for_backward(x8.data, pre8.data, hid8.data, W1, W2, slope.detach(), kd, mask)
        ctx.kind, ctx.D, ctx.Cm, ctx.slope_shape = kind, D, Cm, slope.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import tc
        x8d, pre8d, hid8d, W1, W2, slope, kd, mask = ctx.saved_tensors
        D, Cm, kind = ctx.D, ctx.Cm, ctx.kind
        H = D * Cm
        dy8, _, dbD = tc.dy_prep(dy, None, kind, want_bias=True)
        d_hid8 = tc.conv_tc(dy8, tc.PackedConv(_dgrad_weights(W2), None, kind))
        g8, stats = tc.prelu_bwd_c8(d_hid8, tc.C8(pre8d, H, kind), slope)
        db1 = stats[:H].view(D, Cm).sum(0)
        dslope = stats[g8.Cp:g8.Cp + H].sum().reshape(ctx.slope_shape)
        dW2 = conv2d_wgrad_tc(tc.C8(hid8d, H, kind), dy8, H, D, 3)                      # (D, D*Cm, 3, 3)
        dx = tc.conv_tc(g8, tc.PackedConv(_dgrad_weights(W1), None, kind), out_nchw=True)
        dW1 = conv2d_wgrad_tc(tc.C8(x8d, D, kind), g8, D, H, 3)                         # (D*Cm, D, 3, 3)
        # adjoint of the banded gather: depth tap kd collects the (kd - 1)-th diagonal of the (d, d') block -- plain sums, so the
        # result is bit-reproducible (index_add_ accumulates with atomics in arbitrary order)
        dg1 = dW1.reshape(D, Cm, D, 3, 3).permute(1, 3, 4, 0, 2)                        # [c, h, w, d, d']
        dg2 = dW2.reshape(D, D, Cm, 3, 3).permute(2, 3, 4, 0, 1)
        dw1 = torch.stack([dg1.diagonal(offset=k - 1, dim1=3, dim2=4).sum(-1) for k in range(3)], dim=-1).unsqueeze(1)
        dw2 = torch.stack([dg2.diagonal(offset=k - 1, dim1=3, dim2=4).sum(-1) for k in range(3)], dim=-1).unsqueeze(0)
        return dx, dw1.contiguous(), db1, dslope, dw2.contiguous(), dbD.sum().reshape(1), None


# CWFA_STENCIL_NODE=0 keeps the generic conv2d -> PReLU -> conv2d node chain (A/B measurements)
_STENCIL_NODE = __import__("os").environ.get("CWFA_STENCIL_NODE", "1") == "1"


def depth_stencil3d(x, w1, b1, slope, w2, b2):
    if _PRECISION != "fp32":
        if _STENCIL_NODE and slope.numel() == 1 and w1.shape[1] == 1 and tuple(w1.shape[2:]) == (3, 3, 3) and b1 is not None and b2 is not None:
            return _StencilBandedTC.apply(x, w1, b1, slope, w2, b2, _PRECISION)
        return depth_stencil3d_banded(x, w1, b1, slope, w2, b2)
    return _DepthStencil.apply(x, w1, b1, slope, w2, b2)


# ---------------------------------------------------------------------------------------------
# LRNN U-Net ops (unet.py:72-113,161-195): BatchNorm2d, 2x2 max-pool, ConvTranspose2d(k=2,s=2) + skip
# ---------------------------------------------------------------------------------------------
class _BatchNorm(_F):
    """y = gamma * (x - mean) * rstd + beta with batch statistics (the reference's LRNN runs in .train() mode, CWFA.py:531-532) or
    running statistics.  Adjoint (batch): dx = a dy + b x + c0 with a = gamma rstd, b = -gamma rstd^2 <dy, xhat>/n,
    c0 = -a <dy>/n - b mean; d gamma = <dy, xhat>, d beta = <dy> -- two per-channel sums + one element-wise pass."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, batch_stats, eps, momentum=None, update_running=False,
                num_batches_tracked=None):
        xx = _f32(x)
        N, C = xx.shape[0], xx.shape[1]
        P = xx[0, 0].numel()
        n = float(N * P)
        if batch_stats:
            st = ops.channel_stats(xx)
            if update_running:
                ops.update_running_stats(st.reshape(-1), n, running_mean, running_var, momentum, num_batches_tracked)
            s, q = st.double()
            mean = s / n
            var = (q / n - mean * mean).clamp_min(0.0)
        else:
            mean, var = running_mean.detach().double(), running_var.detach().double()
        rstd = 1.0 / torch.sqrt(var + eps)
        g = gamma.detach().double()
        scale = (g * rstd).float()
        shift = (beta.detach().double() - mean * g * rstd).float()
        ctx.save_for_backward(xx, g, mean, rstd)
        ctx.batch_stats, ctx.n = bool(batch_stats), n
        return ops.scale_shift(xx, scale, shift)

    @staticmethod
    def backward(ctx, dy):
        x, g, mean, rstd = ctx.saved_tensors
        dy = ops._ck(dy)
        N, C = x.shape[0], x.shape[1]
        P = x[0, 0].numel()
        lib = _lib.load()
        out = torch.empty(2 * C, device=x.device, dtype=torch.float32)
        ws = torch.empty(2 * C * lib.cwfa_channel_dot_workspace_blocks(), device=x.device, dtype=torch.float32)
        _lib.call("cwfa_channel_dot_stats_f32", x.data_ptr(), dy.data_ptr(), out.data_ptr(), ws.data_ptr(), N, C, P, _stream())
        sdy, sdyx = out[:C].double(), out[C:].double()
        dgamma = rstd * (sdyx - mean * sdy)
        a = g * rstd
        if ctx.batch_stats:
            b = -g * rstd * rstd * dgamma / ctx.n
            c0 = -a * sdy / ctx.n - b * mean
        else:
            b = torch.zeros_like(a)
            c0 = torch.zeros_like(a)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            af, bf, cf = a.float().contiguous(), b.float().contiguous(), c0.float().contiguous()
            _lib.call("cwfa_bn_bwd_apply_f32", dy.data_ptr(), x.data_ptr(), af.data_ptr(), bf.data_ptr(), cf.data_ptr(), dx.data_ptr(),
                      N, C, P, _stream())
        return dx, dgamma.float(), sdy.float(), None, None, None, None, None, None, None


def batchnorm(x, gamma, beta, running_mean, running_var, *, batch_stats, eps, momentum=None, update_running=False,
              num_batches_tracked=None):
    return _BatchNorm.apply(x, gamma, beta, running_mean, running_var, batch_stats, eps, momentum, update_running, num_batches_tracked)


class _ScaleShift(_F):
    """y[b,c] = x[b,c] * scale[c] + shift[c] (ActNorm, invertible_resnet.py:66-81; global affine of AllInOneBlock,
    all_in_one_block.py:171-187).  Adjoint: dx = dy * scale[c]; d scale[c] = <dy, x>_c; d shift[c] = <dy>_c -- the two
    per-channel sums come from ONE pass of the BatchNorm-adjoint reduction kernel (``channel_dot_stats``)."""

    @staticmethod
    def forward(ctx, x, scale, shift):
        xx, sc, sh = _f32(x), _f32(scale).reshape(-1), _f32(shift).reshape(-1)
        ctx.save_for_backward(xx, sc)
        ctx.shapes = (scale.shape, shift.shape)
        return ops.scale_shift(xx, sc, sh)

    @staticmethod
    def backward(ctx, dy):
        x, sc = ctx.saved_tensors
        dy = ops._ck(dy)
        N, C = x.shape[0], x.shape[1]
        P = x[0, 0].numel()
        dx = ops.scale_shift(dy, sc, torch.zeros_like(sc)) if ctx.needs_input_grad[0] else None
        dsc = dsh = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            out = torch.empty(2 * C, device=x.device, dtype=torch.float32)
            ws = torch.empty(2 * C * _lib.load().cwfa_channel_dot_workspace_blocks(), device=x.device, dtype=torch.float32)
            _lib.call("cwfa_channel_dot_stats_f32", x.data_ptr(), dy.data_ptr(), out.data_ptr(), ws.data_ptr(), N, C, P, _stream())
            dsh, dsc = out[:C].reshape(ctx.shapes[1]), out[C:].reshape(ctx.shapes[0])
        return dx, dsc, dsh


def scale_shift(x, scale, shift):
    return _ScaleShift.apply(x, scale, shift)


class _MaxPool2(_F):
    @staticmethod
    def forward(ctx, x):
        xx = _f32(x)
        ctx.save_for_backward(xx)
        return ops.maxpool2(xx)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = ops._ck(dy)
        N, C, H, W = x.shape
        dx = torch.empty_like(x)
        _lib.call("cwfa_maxpool2_bwd_f32", x.data_ptr(), dy.data_ptr(), dx.data_ptr(), N, C, H, W, _stream())
        return dx


def maxpool2(x):
    return _MaxPool2.apply(x)


class _PixelShuffle2(_F):
    """y[n,c,2h+i,2w+j] = z[n,4c+2i+j,h,w] + skip: the scatter half of ConvTranspose2d(k=2,s=2) (unet.py:166) and the skip ADD
    (unet.py:190); the channel-mixing half is an ordinary 1x1 convolution to 4*Cout channels."""

    @staticmethod
    def forward(ctx, z, skip):
        zz = _f32(z)
        sk = None if skip is None else _f32(skip)
        N, C4, H, W = zz.shape
        y = torch.empty((N, C4 // 4, 2 * H, 2 * W), device=zz.device, dtype=torch.float32)
        _lib.call("cwfa_pixel_shuffle2_f32", zz.data_ptr(), ops._p(sk), y.data_ptr(), N, C4 // 4, H, W, 0, _stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = ops._ck(dy)
        N, C, H2, W2 = dy.shape
        dz = None
        if ctx.needs_input_grad[0]:
            dz = torch.empty((N, 4 * C, H2 // 2, W2 // 2), device=dy.device, dtype=torch.float32)
            _lib.call("cwfa_pixel_shuffle2_f32", dy.data_ptr(), None, dz.data_ptr(), N, C, H2 // 2, W2 // 2, 1, _stream())
        return dz, (dy if ctx.needs_input_grad[1] else None)


def conv_transpose2x2(x, w, bias=None, skip=None):
    """ConvTranspose2d(k=2,s=2)(x) + skip, differentiable: 1x1 convolution with W'[(co,i,j), ci] = w[ci,co,i,j] (a re-indexing of
    the parameter; forward / data gradient / weight gradient on the convolution kernels) followed by the pixel shuffle."""
    Cin, Cout = w.shape[0], w.shape[1]
    wp = w.permute(1, 2, 3, 0).reshape(4 * Cout, Cin, 1, 1)
    bp = None if bias is None else bias.repeat_interleave(4)
    return _PixelShuffle2.apply(conv2d(x, wp, bp), skip)

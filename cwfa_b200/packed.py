"""Packed tensor-core executors of the sub-networks of the path: weights of a module packed ONCE into the layout the tcgen05
kernels stream (cwfa_b200/tc.py), activations in the C8 half-precision layout.

Two users:
* ``cwfa_b200.engine.CWFAEngine`` (whole-frame throughput engine, CUDA-graph replay);
* the drop-in module API itself: with ``cwfa_b200.set_inference_precision('bf16' | 'fp16')`` the ``forward`` of
  ``wavelet_flow_subnetwork2D(_first)``, ``cond_network``, ``Encoder`` and of the coupling blocks routes through these executors
  whenever no gradient is required, so ``conv_inn[n]([z, lo], c=..., rev=True)`` / ``cond_nets[n](views)`` -- the reference's own
  entry points (CWFA.py:181,895,912,966) -- run on the tensor cores.  ``'fp32'`` (default) keeps the reference-precision kernels.

Cache: one executor per (module, kind), rebuilt when a parameter's version / storage changes or after ``weights_changed()``
(called by ``cwfa_b200.training.Lion.step``, whose kernel updates the flat parameter buffer in place).
"""
from __future__ import annotations

import contextlib
import math
from typing import Optional

import torch

from . import ops, tc

_INFER = "fp32"
_EPOCH = 0


def set_inference_precision(kind: str) -> str:
    """'fp32' (reference-precision CUDA-core kernels), 'bf16' or 'fp16' (tcgen05 convolutions with half operands, fp32
    accumulation; couplings, log-dets, Haar stay fp32).  Returns the previous setting."""
    global _INFER
    if kind not in ("fp32", "bf16", "fp16"):
        raise ValueError(kind)
    prev, _INFER = _INFER, kind
    return prev


def inference_precision_kind() -> str:
    return _INFER


@contextlib.contextmanager
def inference_precision(kind: str):
    prev = set_inference_precision(kind)
    try:
        yield
    finally:
        set_inference_precision(prev)


def weights_changed() -> None:
    """Invalidate every cached executor (parameters were modified behind torch's back, e.g. by the Lion kernel)."""
    global _EPOCH
    _EPOCH += 1


def fast_kind(*tensors) -> Optional[str]:
    """The tensor-core kind to use for an inference call, or None: the reference-precision path runs when the switch is
    'fp32' or when autograd needs the call on its tape."""
    if _INFER == "fp32":
        return None
    if torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in tensors):
        return None
    return _INFER


def executor(module: torch.nn.Module, kind: str, factory):
    """Cached ``factory(module, kind)``; the key covers every parameter's in-place version and storage address."""
    params = list(module.parameters()) + list(module.buffers())
    key = (kind, _EPOCH, bool(module.training), tuple((p._version, p.data_ptr()) for p in params))
    cache = module.__dict__.setdefault("_cwfa_exec", {})
    hit = cache.get(kind)
    if hit is None or hit[0] != key:
        hit = cache[kind] = (key, factory(module, kind))
    return hit[1]


class _Subnet:
    """Packed weights of one wavelet_flow_subnetwork2D(_first) (networks.py:586-706)."""

    def __init__(self, sub, kind):
        self.normal = sub.normal
        self.kind = kind
        first = sub.block12 if sub.normal else sub.block1
        self.inp = tc.PackedConv(first.weight, first.bias, kind)
        self.res = []
        for name in ("block2", "block4", "block6"):
            blk = getattr(sub, name)
            self.res.append((tc.PackedConv(blk[0].weight, blk[0].bias, kind), tc.PackedConv(blk[2].weight, blk[2].bias, kind)))
        last = sub.block72[1] if sub.normal else sub.block7[1]
        self.out = tc.PackedConv(last.weight, last.bias, kind, bn=tc.pad16(last.weight.shape[0]))   # one N block: s and t in one CTA

    def trunk(self, lf8: tc.C8, b: Optional[tc.C8] = None, chunk_off: int = 0) -> tc.C8:
        """LF condition (C8) -> ELU(b6) (C8, n hidden channels).  ``b`` (optional): precomputed output of the input
        1x1 conv, possibly a slice (``chunk_off``) of the tensor produced by one batched conv over all
        sub-networks of the level."""
        if b is None:
            b = tc.conv_tc(lf8, self.inp)
        fused = self.inp.Cout_p == 64
        for i, (p3, p1) in enumerate(self.res):
            if fused:
                b = tc.resblock_tc(b, p3, p1, chunk_off if i == 0 else 0)
            else:
                assert chunk_off == 0
                t = tc.conv_tc(b, p3, act=ops.ACT_ELU)
                b = tc.conv_tc(t, p1, act=ops.ACT_ELU, res=b, res_mode=1)
        return b

    def __call__(self, lf8: tc.C8) -> torch.Tensor:
        """LF condition -> fp32 NCHW coefficient tensor (unfused path)."""
        return tc.conv_tc(self.trunk(lf8), self.out, out_nchw=True)

    # ---- module-API fast path -----------------------------------------------------------------------------------------
    def from_nchw(self, inp: torch.Tensor) -> torch.Tensor:
        """``subnet(inp)`` of the reference API: fp32 NCHW in, fp32 NCHW [s | t] out (the ``_first`` variant receives only its
        LF half and returns s)."""
        return self(tc.to_c8(inp, self.kind))

    def couple(self, inp: torch.Tensor, x: Optional[torch.Tensor], *, ch: int, inverse: bool, clamp: float, k_atan: float,
               t_ext: Optional[torch.Tensor] = None, t_scale: float = 1.0):
        """Trunk + last conv with the affine coupling fused into its epilogue (``tc.conv_tc_coupling``): returns (y, logdet[B])."""
        b8 = self.trunk(tc.to_c8(inp, self.kind))
        buf = torch.zeros(b8.N + 1, device=inp.device, dtype=torch.float32)          # log-det accumulator + the kernel's ticket
        y = tc.conv_tc_coupling(b8, self.out, x, ch=ch, inverse=inverse, clamp=clamp, k_atan=k_atan, t_ext=t_ext, t_scale=t_scale,
                                logdet=buf[:b8.N], ticket=buf[b8.N:].view(torch.int32))
        return y, buf[:b8.N]


class _CondNet:
    """cond_network / ResidualBlock (networks.py:165-242) entirely on the tensor cores.

    The depth stencil Conv3d(1->Cm) -> PReLU -> Conv3d(Cm->1) over the (H, W, depth) volume
    (networks.py:221-225,239) runs as the fused voxel-row kernel ``tc.stencil3d_tc`` for the reference's Cm = 32; for any other
    width it is executed as two ordinary 3x3 2-D convolutions whose channel axis
    carries (depth, hidden-channel) and whose weights are the depth-banded expansion of the 3x3x3
    kernels: W1[(d,c), d'] = w1[c, :, :, d'-d+1], W2[d, (d',c)] = w2[c, :, :, d'-d+1] for |d'-d| <= 1.
    Zero padding in depth falls out of the band; zero padding in H, W is the conv's own padding."""

    def __init__(self, net, kind):
        if getattr(net, "global_attention", None) is not None:
            raise NotImplementedError("the tensor-core conditioning-net executor takes the views as they are: apply "
                                      "cond_network.global_attention (None in the reference, networks.py:189) in the module path")
        rb = net.subnetworks[0]
        self.rb = rb
        self.c1 = tc.PackedConv(rb.conv1[0].weight, rb.conv1[0].bias, kind)
        self.ds = tc.PackedConv(rb.downsample[0].weight, rb.downsample[0].bias, kind)
        self.c2 = tc.PackedConv(rb.conv2[0].weight, rb.conv2[0].bias, kind)
        w1, b1 = rb.conv3d[0].weight.detach().float(), rb.conv3d[0].bias.detach().float()      # (Cm,1,3,3,3), (Cm)
        w2, b2 = rb.conv3d[3].weight.detach().float(), rb.conv3d[3].bias.detach().float()      # (1,Cm,3,3,3), (1)
        Cm, D = w1.shape[0], rb.out_channels
        dev = w1.device
        self.D = D
        # The fused true-3-D kernel (csrc/stencil_tc.cu: voxel rows, hidden volume kept in the SM) covers the reference's
        # default 32 hidden channels; other widths take the banded two-convolution form below.
        self.fused = Cm == 32 and 1 <= D <= 64 and rb.conv3d[1].weight.numel() == 1 and self.fuse_stencil
        if self.fused:
            self.sw = tc.StencilWeights(w1, b1, w2, b2, kind)
            return
        W1 = torch.zeros(D, Cm, D, 3, 3, device=dev)          # [d, c, d', ky, kx]
        W2 = torch.zeros(D, D, Cm, 3, 3, device=dev)          # [d, d', c, ky, kx]
        for kd in range(3):
            for d in range(D):
                dp = d + kd - 1
                if 0 <= dp < D:
                    W1[d, :, dp] = w1[:, 0, :, :, kd]
                    W2[d, dp] = w2[0, :, :, :, kd]
        self.s1 = tc.PackedConv(W1.reshape(D * Cm, D, 3, 3), b1.repeat(D), kind)
        # second stencil conv (Cin = 32 D, Cout = D): evaluated as ONE 1x1 conv to 9 tap partials per depth + a col2im
        # sum -- the tap-by-tap form re-reads the 32 D-channel operand tile from shared memory nine times for a tiny N.
        # n-blocks of 144 channels = 16 output depths: the packer's zero K-block masks skip the hidden depths out of reach.
        Wg = tc.col2im3x3_weights(W2.reshape(D, D * Cm, 3, 3))
        gp = tc.pad16(Wg.shape[0])
        self.s2g = tc.PackedConv(Wg, None, kind, bn=144 if gp % 144 == 0 else gp)
        self.s2_bias = torch.zeros(tc.pad16(D), device=dev, dtype=torch.float32)
        self.s2_bias[:D] = b2
        self.s2_mb = 1 if self.s2g.BN == 144 else 2       # measured (scripts/bench_stencil.py)

    # CWFA_FUSED_STENCIL=0 keeps the banded form (A/B measurements)
    fuse_stencil = __import__("os").environ.get("CWFA_FUSED_STENCIL", "1") == "1"

    def __call__(self, v8: tc.C8) -> tc.C8:
        """views (C8) -> LF condition (C8)."""
        rb = self.rb
        out = tc.conv_tc(v8, self.c1, act=ops.ACT_PRELU, slope=rb.conv1[1].weight)
        res = tc.conv_tc(v8, self.ds)
        out = tc.conv_tc(out, self.c2, act=ops.ACT_PRELU, slope=rb.relu.weight, res=res, res_mode=1)
        if self.fused:
            return tc.stencil3d_tc(out, self.sw, rb.conv3d[1].weight, self.D)
        hid = tc.conv_tc(out, self.s1, act=ops.ACT_PRELU, slope=rb.conv3d[1].weight)
        return tc.col2im3x3_c8(tc.conv_tc(hid, self.s2g, mb=self.s2_mb), self.s2_bias, self.D)

    def from_nchw(self, views: torch.Tensor) -> torch.Tensor:
        """``cond_net(views)[-1]`` of the reference API: fp32 NCHW in and out."""
        return tc.from_c8(self(tc.to_c8(views, self.c1.kind)))


class _UNet:
    """LRNN U-Net (unet.py:9-195) in C8: conv+PReLU on tensor cores, BatchNorm (+max-pool) fused passes."""

    def __init__(self, unet, kind):
        self.unet = unet

        def block(b):
            return [(tc.PackedConv(b.block[i].weight, b.block[i].bias, kind), b.block[i + 1], b.block[i + 2]) for i in (0, 3)]

        self.down = [block(d) for d in unet.down_path]
        # transposed convs: the deepest one (few tiles: 1.7 waves of 256-column items) runs better as 128-column,
        # double-buffered items (measured: 131 -> 112 us, scripts/bench_convT.py)
        self.up = [(tc.PackedConv(u.up.weight, u.up.bias, kind, transposed=True, bn=128 if u.up.weight.shape[0] >= 1024 else None),
                    block(u.conv_block)) for u in unet.up_path]
        self.last = tc.PackedConv(unet.last[0].weight, unet.last[0].bias, kind)

    # BatchNorm batch statistics produced by the conv's epilogue (cwfa_conv_tc_bn) instead of a separate pass over the tensor.
    # Measured NEUTRAL at frame level on B200 (A/B in one run: 153.4 / 154.7 frames/s fused vs 155.6 / 153.5 separate): the
    # butterfly + partial stores cost the 256-channel convs what the saved 0.9 GB read pass gains, so the default keeps the
    # convolution kernels lean; CWFA_FUSE_BN=1 (or the class attribute) switches it on.
    fuse_bn_stats = __import__("os").environ.get("CWFA_FUSE_BN", "0") == "1"

    def _conv_bn(self, x, pconv, prelu, bn, training, pool=False):
        """conv -> PReLU -> BatchNorm (unet.py:99-107).  In batch-statistics mode the per-channel sums come out of the conv's own
        epilogue (no pass over the tensor for them); the normalisation (+ the 2x2 max-pool) is one apply pass."""
        if training and self.fuse_bn_stats:
            y, part, mb = tc.conv_tc_bn_stats(x, pconv, act=ops.ACT_PRELU, slope=prelu.weight)
            return tc.batchnorm_c8(y, bn.weight, bn.bias, bn.running_mean, bn.running_var, batch_stats=True, eps=bn.eps, pool=pool,
                                   partial=(part, mb))
        y = tc.conv_tc(x, pconv, act=ops.ACT_PRELU, slope=prelu.weight)
        return tc.batchnorm_c8(y, bn.weight, bn.bias, bn.running_mean, bn.running_var, batch_stats=training, eps=bn.eps, pool=pool)

    def _block(self, x, blk, training, pool):
        (p0, a0, n0), (p1, a1, n1) = blk
        x = self._conv_bn(x, p0, a0, n0, training)
        return self._conv_bn(x, p1, a1, n1, training, pool=pool)

    def __call__(self, x8: tc.C8) -> torch.Tensor:
        training = self.unet.training
        skips = []
        nd = len(self.down)
        for i, blk in enumerate(self.down):
            if i != nd - 1:
                full, x8 = self._block(x8, blk, training, True)
                skips.append(full)
            else:
                x8 = self._block(x8, blk, training, False)
        for i, (up, blk) in enumerate(self.up):
            x8 = tc.conv_transpose_tc(x8, up, skips[-i - 1], mb=1 if up.BN == 128 else None)
            x8 = self._block(x8, blk, training, False)
        return tc.conv_tc(x8, self.last, act=ops.ACT_PRELU, slope=self.unet.last[1].weight, out_nchw=True)


class _LRNN:
    """Encoder/LRNN (networks.py:505-584)."""

    def __init__(self, enc, kind):
        net = enc.net
        self.net = net
        self.kind = kind
        self.proj = tc.PackedConv(net.deconv[0].weight, net.deconv[0].bias, kind)
        self.unet = _UNet(net.deconv[1], kind)
        cn0, cn1 = net.conv3d[0], net.conv3d[1]
        self.cn0_in = tc.PackedConv(cn0.input.weight, cn0.input.bias, kind)          # 1x1  6 -> 64
        self.cn0_7x7 = tc.PackedConv(cn0.m[0].weight, cn0.m[0].bias, kind)           # 7x7 64 -> 64
        self.cn0_1x1 = tc.PackedConv(cn0.m[2].weight, cn0.m[2].bias, kind)           # 1x1 64 -> 64
        self.cn1_in = tc.PackedConv(cn1.input.weight, cn1.input.bias, kind)          # 1x1 64 -> 6
        self.cn1_7x7 = tc.PackedConv(cn1.m[0].weight, cn1.m[0].bias, kind)           # 7x7  6 -> 6
        # element-wise LayerNorm parameters of the wide ConvNeXt block in the activation layout (half the traffic)
        self.cn0_ln_w = tc.to_c8(cn0.m[1].weight.detach().float()[None].contiguous(), kind)
        self.cn0_ln_b = tc.to_c8(cn0.m[1].bias.detach().float()[None].contiguous(), kind)

    def _mean_branch(self, mean_vol: torch.Tensor) -> torch.Tensor:
        """conv3d = ConvNeXt(6,64) -> ConvNeXt(64,6) on the mean volume (networks.py:486-503,527-530): every conv on
        the tensor cores (channels padded to 16); the 64-channel LayerNorm([C,H,W]) runs on the C8 tensor, the 6-channel
        one and the final 6-channel 1x1 stay fp32."""
        cn0, cn1 = self.net.conv3d[0], self.net.conv3d[1]
        k = self.kind
        up8 = tc.conv_tc(tc.to_c8(mean_vol, k), self.cn0_in)                                        # C8, 64 ch
        m8 = tc.layernorm_c8(tc.conv_tc(up8, self.cn0_7x7), self.cn0_ln_w, self.cn0_ln_b, cn0.m[1].eps)
        y8 = tc.conv_tc(m8, self.cn0_1x1, act=ops.ACT_GELU, res=up8, res_mode=2)                    # GELU(.) + up
        up1_8 = tc.conv_tc(y8, self.cn1_in)                                                         # C8, 6 (16) ch
        m1 = tc.conv_tc(up1_8, self.cn1_7x7, out_nchw=True)
        m1 = ops.layernorm_chw(m1, cn1.m[1].weight, cn1.m[1].bias, cn1.m[1].eps)
        return ops.conv2d(m1, cn1.m[2].weight, cn1.m[2].bias, act=ops.ACT_GELU, res=tc.from_c8(up1_8), res_mode=2)

    def __call__(self, v8: tc.C8, mean_vol: Optional[torch.Tensor]) -> torch.Tensor:
        x = self.unet(tc.conv_tc(v8, self.proj))
        if mean_vol is not None:
            x = ops.attention_gate_(x, self._mean_branch(mean_vol), mean_vol, self.net.attention_3d)
        return x

    def from_nchw(self, views: torch.Tensor, mean_vol: Optional[torch.Tensor]) -> torch.Tensor:
        return self(tc.to_c8(views, self.kind), mean_vol)



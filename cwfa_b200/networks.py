"""Model construction of the CWFA hot path, mirroring the reference's ``networks.py`` /
``unet.py`` public names, constructor arguments and ``state_dict`` layout.

The ``nn.Conv2d`` / ``nn.BatchNorm2d`` / ... children below are PARAMETER HOLDERS only (same
key names, shapes and default initialisation as the reference); their own ``forward`` is
never called.  Every ``forward`` here launches this repo's CUDA kernels through ``ops``
(reference precision fp32).  With gradients enabled the same calls go on torch's autograd tape
(``cwfa_b200.autograd``: forward and adjoint kernels, optionally on the tcgen05 convolution kernels via
``autograd.set_training_precision('bf16'|'fp16')``); the throughput path for inference is ``cwfa_b200.engine``.

Inference semantics: stochastic regularisers are identity (Dropout3d of the conditioning net
in eval mode, networks.py:224; the U-Net's always-on ``F.dropout2d(p=0.005)``, unet.py:80,86;
ConvNeXt drop_path, networks.py:502).  BatchNorm uses batch statistics when the module is in
``.train()`` mode (what the reference runs, CWFA.py:531-532) and running statistics in
``.eval()`` mode.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import framework as Ff
from . import modules as Fm
from . import ops, packed
from .modules import HaarTransform1D, PermuteDim

# hidden width of the coupling sub-networks; set by conditional_wavelet_flow (networks.py:272-274)
networks_n_chans = 64


# ---------------------------------------------------------------------------------------------
# initialisers (networks.py:19-96)
# ---------------------------------------------------------------------------------------------
def subnet_initialization_small(m):
    if isinstance(m, (nn.Conv2d, nn.Conv3d, nn.Linear)):
        nn.init.xavier_uniform_(m.weight.data, 0.01)
        if m.bias is not None:
            m.bias.data *= 0.01


def subnet_initialization_positive(m):
    if isinstance(m, (nn.Conv2d, nn.Conv3d, nn.Linear)):
        nn.init.xavier_uniform_(m.weight.data, 0.1)
        m.weight.data = m.weight.data.abs()
        if m.bias is not None:
            m.bias.data *= 0.1


def _conv(holder: nn.Module, x, **kw):
    return ops.conv2d(x, holder.weight, holder.bias, **kw)


# ---------------------------------------------------------------------------------------------
# coupling sub-network (networks.py:586-706)
# ---------------------------------------------------------------------------------------------
class wavelet_flow_subnetwork2D(nn.Module):
    """1x1 c_in->n ; 3 x [3x3 n->n, ELU, 1x1 n->n, +res, (ELU)] ; ELU ; 3x3 n->c_out.
    Reference: wavelet_flow_subnetwork.forward networks.py:641-671, init_blocks :608-638."""

    normal = True

    def __init__(self, c_in, c_out, c_internal=[]):
        super().__init__()
        self.c_in, self.c_out, self.c_internal = c_in, c_out, c_internal
        n = self.n_ch = networks_n_chans
        self.block_grad_up = nn.Conv2d(c_in // 2, c_in, 3, padding=1)          # unused, kept for checkpoints
        self.block1 = nn.Conv2d(c_in // 2, n, 1)
        self.block12 = nn.Conv2d(c_in, n, 1)
        for name in ("block2", "block4", "block6"):
            setattr(self, name, nn.Sequential(nn.Conv2d(n, n, 3, 1, 1), nn.ELU(), nn.Conv2d(n, n, 1)))
        self.block3, self.block5 = nn.ELU(), nn.ELU()
        self.block7 = nn.Sequential(nn.ELU(), nn.Conv2d(n, c_out // 2, 3, 1, 1))
        self.block72 = nn.Sequential(nn.ELU(), nn.Conv2d(n, c_out, 3, 1, 1))

    def trunk(self, b1):
        """b1 -> ELU(b6): every ELU is fused into the producing conv's epilogue."""
        b = b1
        for i, name in enumerate(("block2", "block4", "block6")):
            blk = getattr(self, name)
            t = _conv(blk[0], b, act=ops.ACT_ELU)
            # b_{k+1} = ELU(conv1x1(t) + b_k); the ELU after block6 belongs to block7/72 (networks.py:635-638)
            b = _conv(blk[2], t, res=b, res_mode=1, act=ops.ACT_ELU)
        return b

    def packed_executor(self, kind: str):
        """Tensor-core executor of this sub-network (weights packed once, cached; ``cwfa_b200.packed``)."""
        return packed.executor(self, kind, packed._Subnet)

    def forward(self, inp):
        kind = packed.fast_kind(inp)
        if kind is not None:                     # inference with set_inference_precision('bf16'|'fp16'): tcgen05 kernels
            return self.packed_executor(kind).from_nchw(inp)
        return self._train_path(inp, self.block12, self.block72[1])

    def _train_path(self, inp, conv_in, conv_out):
        """conv_out(trunk(conv_in(inp))) with gradients: one C8-native autograd node under a half training precision
        (``autograd._SubnetTC``), else one node per convolution."""
        from . import autograd as ag
        blocks = [(getattr(self, nm)[0], getattr(self, nm)[2]) for nm in ("block2", "block4", "block6")]
        convs = [conv_in, conv_out] + [c for b in blocks for c in b]
        if ag.subnet_tc_supported(inp, convs):
            return ag.subnet_tc(inp, conv_in, blocks, conv_out)
        return _conv(conv_out, self.trunk(_conv(conv_in, inp)))


class wavelet_flow_subnetwork2D_first(wavelet_flow_subnetwork2D):
    """First block of a level: input = cat(meanvol_delta, LF).  s = trunk(LF) through block7,
    t = -meanvol_delta / sqrt2 (networks.py:653-657, :669-671)."""

    normal = False

    def __init__(self, c_in, c_out, c_internal=[]):
        super().__init__(c_in, c_out, c_internal)
        self.block7[-1].apply(subnet_initialization_small)

    def forward_split(self, inp):
        """Returns (a_s, a_t, t_scale) without materialising cat(b7, -low/sqrt2)."""
        n = self.c_in // 2
        low, cond = inp[:, :-n], inp[:, -n:]
        kind = packed.fast_kind(inp)
        if kind is not None:
            return self.packed_executor(kind).from_nchw(cond.contiguous()), low, -1.0 / math.sqrt(2)
        b7 = self._train_path(cond.contiguous(), self.block1, self.block7[1])
        return b7, low, -1.0 / math.sqrt(2)

    def forward(self, inp):
        b7, low, ts = self.forward_split(inp)
        return torch.cat((b7, low * ts), 1)


# the 3-D base class of the reference is never used by CWFA; the name is kept as an alias
wavelet_flow_subnetwork = wavelet_flow_subnetwork2D


# ---------------------------------------------------------------------------------------------
# conditioning network (networks.py:165-242)
# ---------------------------------------------------------------------------------------------
_SHARED_PRELU = nn.PReLU()     # the reference shares ONE default-arg instance (networks.py:209)


class ResidualBlock(nn.Module):
    def __init__(self, in_channels, out_channels, chans_3D=32, stride=1, downsample=None, activation=_SHARED_PRELU):
        super().__init__()
        if stride != 1:
            raise ValueError("cwfa_b200.ResidualBlock supports stride 1 only (the only value CWFA uses)")
        self.conv1 = nn.Sequential(nn.Conv2d(in_channels, out_channels, 3, stride, 1), activation)
        self.conv2 = nn.Sequential(nn.Conv2d(out_channels, out_channels, 3, 1, 1))
        self.downsample = nn.Sequential(nn.Conv2d(in_channels, out_channels, 3, stride, 1))
        self.relu = activation
        self.conv3d = nn.Sequential(nn.Conv3d(1, chans_3D, 3, stride, 1), activation, nn.Dropout3d(),
                                    nn.Conv3d(chans_3D, 1, 3, stride, 1))
        self.bn_out = None
        self.out_channels = out_channels

    def forward(self, x):
        """PReLU(conv2(PReLU(conv1 x)) + downsample(x)) -> depth stencil.  networks.py:229-242."""
        a = self.relu.weight
        out = _conv(self.conv1[0], x, act=ops.ACT_PRELU, slope=self.conv1[1].weight)
        res = _conv(self.downsample[0], x)
        out = _conv(self.conv2[0], out, res=res, res_mode=1, act=ops.ACT_PRELU, slope=a)
        c3 = self.conv3d
        return ops.depth_stencil3d(out, c3[0].weight, c3[0].bias, c3[1].weight, c3[3].weight, c3[3].bias)


class cond_network(nn.Module):
    """29 lenslet views -> per-level LF condition.  Returns a one-element list (networks.py:195-196)."""

    def __init__(self, c_in, c_out, n_steps, max_steps=7, n_channels=[], cond_chans=32, net_constructor=None):
        super().__init__()
        self.n_steps = n_steps
        self.global_attention = None
        self.subnetworks = nn.Sequential(ResidualBlock(c_in, c_out, chans_3D=cond_chans))

    def forward(self, lf_img):
        if self.global_attention is not None:                # networks.py:196 (None in the reference's constructor)
            return [self.subnetworks[0](lf_img * self.global_attention(lf_img))]
        kind = packed.fast_kind(lf_img)
        if kind is not None:
            return [packed.executor(self, kind, packed._CondNet).from_nchw(lf_img)]
        return [self.subnetworks[0](lf_img)]


class GlobalAttention(nn.Module):
    """sigmoid(Conv1d_1(ReLU(Conv1d_3(x flattened over H*W)))).  networks.py:244-262."""

    def __init__(self, n_chans):
        super().__init__()
        self.m = nn.Sequential(nn.Conv1d(n_chans, n_chans, 3, 1, 1), nn.ReLU(), nn.Conv1d(n_chans, n_chans, 1, 1, 0),
                               nn.Sigmoid())

    def forward(self, inp):
        f = inp.reshape(inp.shape[0], inp.shape[1], -1)
        f = ops.conv1d_flat(f, self.m[0].weight, self.m[0].bias, ops.ACT_RELU)
        f = ops.conv1d_flat(f, self.m[2].weight, self.m[2].bias, ops.ACT_SIGMOID)
        return f.reshape(inp.shape)


class ConvNeXt(nn.Module):
    """1x1 -> [7x7, LayerNorm([C,S,S]), 1x1, GELU] + skip.  networks.py:468-503 (drop_path = identity)."""

    def __init__(self, c_in, c_out, drop_prob=0.1, size=512):
        super().__init__()
        self.drop_prob = drop_prob
        self.input = nn.Conv2d(c_in, c_out, 1, 1)
        self.m = nn.Sequential(nn.Conv2d(c_out, c_out, 7, 1, 3), nn.LayerNorm([c_out, size, size]),
                               nn.Conv2d(c_out, c_out, 1, 1), nn.GELU())

    def forward(self, inp):
        up = _conv(self.input, inp)
        m = _conv(self.m[0], up)
        m = ops.layernorm_chw(m, self.m[1].weight, self.m[1].bias, self.m[1].eps)
        return _conv(self.m[2], m, act=ops.ACT_GELU, res=up, res_mode=2)


# ---------------------------------------------------------------------------------------------
# U-Net (unet.py:9-195)
# ---------------------------------------------------------------------------------------------
class UNetConvBlock(nn.Module):
    def __init__(self, in_size, out_size, padding, batch_norm, kernel_size=3, use_bias=False, stride=1,
                 activation=nn.LeakyReLU):
        super().__init__()
        if activation is not nn.PReLU:
            raise ValueError("cwfa_b200.UNet implements the PReLU activation CWFA uses (unet.py:22)")
        block = [nn.Conv2d(in_size, out_size, kernel_size, stride, int(padding), bias=use_bias), activation()]
        if batch_norm:
            block.append(nn.BatchNorm2d(out_size))
        block += [nn.Conv2d(out_size, out_size, kernel_size, 1, int(padding), bias=use_bias), activation()]
        if batch_norm:
            block.append(nn.BatchNorm2d(out_size))
        self.block = nn.Sequential(*block)
        self.batch_norm = batch_norm

    def forward(self, x):
        step = 3 if self.batch_norm else 2
        for i in (0, step):
            x = _conv(self.block[i], x, act=ops.ACT_PRELU, slope=self.block[i + 1].weight)
            if self.batch_norm:
                bn = self.block[i + 2]
                # .train() mode = batch statistics AND the running-statistics update, exactly like nn.BatchNorm2d (the reference
                # keeps its LRNN in .train() mode at inference too, CWFA.py:531-532, so its checkpoints carry these updates)
                x = ops.batchnorm(x, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                  batch_stats=self.training, eps=bn.eps, momentum=bn.momentum,
                                  update_running=self.training and bn.track_running_stats,
                                  num_batches_tracked=bn.num_batches_tracked)
        return x


class UNetUpBlock(nn.Module):
    def __init__(self, in_size, out_size, up_mode, padding, batch_norm, use_bias=False, skip_conn=True,
                 activation=nn.Softplus):
        super().__init__()
        if up_mode != "upconv":
            raise ValueError("cwfa_b200.UNet implements up_mode='upconv' (networks.py:535)")
        self.skip_conn = skip_conn
        self.up = nn.ConvTranspose2d(in_size, out_size, kernel_size=2, stride=2, bias=use_bias)
        in_size = out_size if not skip_conn else in_size // 2
        self.conv_block = UNetConvBlock(in_size, out_size, padding, batch_norm, use_bias=use_bias, activation=activation)

    def forward(self, x, bridge):
        # skip connection is an ADD (unet.py:190); fused into the transposed-conv epilogue
        up = ops.conv_transpose2x2(x, self.up.weight, self.up.bias, bridge if self.skip_conn else None)
        return self.conv_block(up)


class UNet(nn.Module):
    def __init__(self, in_channels=1, n_classes=2, depth=5, wf=6, padding=True, batch_norm=True, up_mode="upsample",
                 drop_out=0, use_bias=False, skip_conn=False, activation=nn.PReLU):
        super().__init__()
        if not padding:
            raise ValueError("cwfa_b200.UNet implements padding=True only")
        self.padding, self.depth, self.skip_conn, self.drop_out = padding, depth, skip_conn, drop_out
        prev = in_channels
        self.down_path = nn.ModuleList()
        for i in range(depth):
            self.down_path.append(UNetConvBlock(prev, 2 ** (wf + i), padding, batch_norm, use_bias=use_bias, activation=activation))
            prev = 2 ** (wf + i)
        self.up_path = nn.ModuleList()
        for i in reversed(range(depth - 1)):
            self.up_path.append(UNetUpBlock(prev, 2 ** (wf + i), up_mode, padding, batch_norm, use_bias=use_bias,
                                            skip_conn=skip_conn, activation=activation))
            prev = 2 ** (wf + i)
        self.last = nn.Sequential(nn.Conv2d(prev, n_classes, kernel_size=1, bias=use_bias), activation())

    def forward(self, x, store_activations=False):
        blocks = []
        for i, down in enumerate(self.down_path):
            x = down(x)
            if i != len(self.down_path) - 1:
                blocks.append(x)
                x = ops.maxpool2(x)          # adaptive_max_pool2d(W/2) == 2x2 max-pool (unet.py:79)
        for i, up in enumerate(self.up_path):
            x = up(x, blocks[-i - 1])
        return _conv(self.last[0], x, act=ops.ACT_PRELU, slope=self.last[1].weight)


# ---------------------------------------------------------------------------------------------
# LRNN / Encoder (networks.py:505-584)
# ---------------------------------------------------------------------------------------------
class LRNN(nn.Module):
    def __init__(self, ch_in, n_depths, use_bias=False, activation=None, size=512):
        super().__init__()
        self.conv3d = nn.Sequential(ConvNeXt(n_depths, 64, 0.05, size=size), ConvNeXt(64, n_depths, 0.05, size=size))
        self.attention_3d = GlobalAttention(n_depths)
        self.deconv = nn.Sequential(
            nn.Conv2d(ch_in, n_depths, 1, stride=1, padding=0, bias=bool(use_bias)),
            UNet(n_depths, n_depths, depth=3, wf=8, drop_out=0.005, use_bias=bool(use_bias), skip_conn=True,
                 up_mode="upconv", batch_norm=True))
        self.deconv[0].apply(subnet_initialization_positive)

    def forward(self, x_in, mean_vol=None):
        x = self.deconv[1](_conv(self.deconv[0], x_in))
        if mean_vol is not None:
            mean_processed = self.conv3d[1](self.conv3d[0](mean_vol))
            if torch.is_grad_enabled() and (mean_processed.requires_grad or self.attention_3d.m[0].weight.requires_grad):
                from . import autograd as ag          # training: differentiable gate, attention through the conv kernels
                x = ag.gate_add(x, mean_processed, self.attention_3d(mean_vol))
            elif mean_vol.shape[1] <= 16:
                x = ops.attention_gate_(x, mean_processed, mean_vol, self.attention_3d)   # x += m*2*(attn-0.5), :554
            else:
                x = ops.gate_add_(x, mean_processed, self.attention_3d(mean_vol))
        return x


class Encoder(nn.Module):
    """Lowest-resolution step: views (+ mean volume) -> (B, n_depths/2^(L-1), S, S).  networks.py:557-584.
    ``size`` (not in the reference signature, default 512) sets the LayerNorm shape of the
    mean-volume branch, which the reference hard-codes to 512 (networks.py:472,490)."""

    def __init__(self, c_in, c_out, n_steps, n_channels=[], use_bias=False, size=512):
        super().__init__()
        self.net = LRNN(c_in, c_out, use_bias, size=size)

    def forward(self, im_in, mean_vol=None):
        kind = packed.fast_kind(im_in, mean_vol)
        if kind is not None and self._tc_ok(mean_vol):
            return [packed.executor(self, kind, packed._LRNN).from_nchw(im_in, mean_vol)]
        return [self.net(im_in) if mean_vol is None else self.net(im_in, mean_vol)]

    def _tc_ok(self, mean_vol) -> bool:
        """The C8 U-Net executor needs unpadded channel counts (multiples of 16: wf = 8 gives 256/512/1024) and the fused
        attention gate its compile-time channel limit."""
        widths_ok = all(blk.block[0].weight.shape[0] % 16 == 0 for blk in self.net.deconv[1].down_path)
        return widths_ok and (mean_vol is None or mean_vol.shape[1] <= 16)


# ---------------------------------------------------------------------------------------------
# flow builder (networks.py:264-368)
# ---------------------------------------------------------------------------------------------
def conditional_wavelet_flow(input_volume_shape, condition_shape, st_subnet, conditional_network, n_down_steps=2,
                             use_permutations=False, block_type="RNVP", n_internal_ch=128, n_blocks=1,
                             disable_low_res_input=False, device="cpu"):
    """Builds ``n_down_steps`` GraphINNs; graph k < last is Haar1D + Split only, the last one also
    carries the conditional flow on the detail half:
    Haar1D -> Split -> CAT_first(c=[meanvol, LF]) -> [Perm -> block(c=LF)] x n_blocks -> PermuteRandom.
    Node order, condition order, seeds (k+nn) and names follow networks.py:305-366 exactly.
    Returns (cond_net, [GraphINN...])."""
    global networks_n_chans
    networks_n_chans = n_internal_ch
    if conditional_network is None:
        cond_net = None
        cond_channels = list(condition_shape[1:])
    else:
        cond_net = conditional_network().to(device)
        # the reference runs the net once on random input to learn the condition shape (:282-283);
        # the conditioning net preserves H,W and emits out_channels, so it is read off directly.
        cond_channels = [cond_net.subnetworks[0].out_channels] + list(condition_shape[2:])

    blocks = {"RNVP": Fm.RNVPCouplingBlock, "GLOW": Fm.GLOWCouplingBlock, "GIN": Fm.GINCouplingBlock,
              "CAT": Fm.ConditionalAffineTransform, "AI1": Fm.AllInOneBlock}
    if block_type not in blocks:
        raise ValueError(f"block_type {block_type!r} is not implemented (have {sorted(blocks)})")
    INN_block = blocks[block_type]

    subnetworks = []
    for k in range(n_down_steps):
        nodes = [Ff.InputNode(*input_volume_shape, name=f"input {k}")]
        nodes.append(Ff.Node(nodes[-1], HaarTransform1D, {"order_by_wavelet": True}, name=f"down_sampling_{k}"))
        n_ch = nodes[-1].output_dims[0][0]
        s0 = int(n_ch * 0.5)
        split1 = Ff.Node(nodes[-1], Fm.Split, {"section_sizes": (s0, n_ch - s0), "dim": 0}, name=f"Split {k}")
        nodes.append(split1)
        last = k == n_down_steps - 1
        if last:
            cshape = list(cond_channels)
            cond = [Ff.ConditionNode(*cshape, name=f"Condition {k-1}")]
            if not disable_low_res_input:
                cond.append(Ff.ConditionNode(*cshape, name=f"Condition I {k-1}"))
                nodes.append(cond[1])
            nodes.append(cond[0])
            first_subnet = wavelet_flow_subnetwork2D if disable_low_res_input else wavelet_flow_subnetwork2D_first
            nodes.append(Ff.Node(split1.out1, Fm.ConditionalAffineTransform, {"subnet_constructor": first_subnet},
                                 conditions=cond, name=f"Block_net{k}_input"))
            for nn_ in range(1, n_blocks + 1):
                nodes.append(Ff.Node(nodes[-1], PermuteDim if nn_ % 2 == 0 else Fm.PermuteRandom, {"seed": k + nn_},
                                     name=f"Permute_net{k}_{nn_}"))
                nodes.append(Ff.Node(nodes[-1], INN_block, {"subnet_constructor": st_subnet}, conditions=[cond[-1]],
                                     name=f"Block_net{k}_{nn_}"))
            if use_permutations:
                nodes.append(Ff.Node(nodes[-1], Fm.PermuteRandom, {}, name="Permute_final2"))
        nodes.append(Ff.OutputNode(nodes[-1] if last else nodes[-1].out1, name=f"Output WVF{k}"))
        nodes.append(Ff.OutputNode(split1.out0, name=f"Output_net{k}"))
        input_volume_shape = split1.output_dims[0]
        subnetworks.append(Ff.GraphINN(nodes))
    return cond_net, subnetworks


def reset_ActNorm(network, n_to_reset=50):
    """Re-arms the data-dependent initialisation of the first ``n_to_reset`` ActNorm layers of an INN (networks.py:137-151; call site
    CWFA.py:537).  Returns (network, number of layers reset)."""
    n = 0
    for m in next(network.named_children())[1]:
        if isinstance(m, Fm.ActNorm):
            m.init_on_next_batch = True
            n += 1
            if n_to_reset and n >= n_to_reset:
                break
    return network, n


def level_spec(inn: "Ff.GraphINN") -> dict:
    """Node sequence of the flow branch of a level (what ``state_dict`` does not carry):
    used by the fused engine and by the parity tests to drive the oracle."""
    nodes = []
    for i, m in enumerate(inn.module_list):
        if isinstance(m, (HaarTransform1D, Fm.Split)):
            continue
        if isinstance(m, Fm.ConditionalAffineTransform):
            nodes.append({"idx": i, "type": "cat" if m.subnet.normal else "cat_first"})
        elif isinstance(m, Fm.PermuteRandom):
            nodes.append({"idx": i, "type": "perm_chan"})
        elif isinstance(m, PermuteDim):
            nodes.append({"idx": i, "type": "perm_dim", "axis": m.axis})
        elif isinstance(m, Fm.GINCouplingBlock):
            nodes.append({"idx": i, "type": "GIN"})
        elif isinstance(m, Fm.GLOWCouplingBlock):
            nodes.append({"idx": i, "type": "GLOW"})
        elif isinstance(m, Fm.RNVPCouplingBlock):
            nodes.append({"idx": i, "type": "RNVP"})
        elif isinstance(m, Fm.AllInOneBlock):
            nodes.append({"idx": i, "type": "AI1"})
        else:
            raise ValueError(f"unsupported module in flow level: {type(m).__name__}")
    return {"nodes": nodes}

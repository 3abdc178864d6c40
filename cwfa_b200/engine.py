"""Throughput engine for the CWFA hot path: the same arithmetic as ``CWFAModel.reconstruct`` /
``forward_nll`` (pipeline.py) but with every wide convolution on the tcgen05 tensor-core kernel
(C8 half-precision activations, fp32 accumulation), weights packed once, no per-call parameter
handling, and optional whole-frame CUDA-graph replay.

Numerics: operands are rounded to ``kind`` ('bf16' default, 'fp16'); the coupling coefficients
(the last conv of every sub-network), the affine couplings, log-dets, the Haar transforms, the
depth stencil of the conditioning net and the 6-channel mean-volume branch stay in fp32.
Stated tolerance vs the fp32 reference: rel-L2 <= 2e-2 (bf16) / 3e-3 (fp16) on the reconstruction
(tests/test_gpu_engine.py reports the measured values per level).

Reference path: CWFA.py:865-924 (inverse), :156-196 / :966-978 (forward NLL).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch

from . import modules as Fm
from . import networks, ops, tc
from .pipeline import CWFAModel, lrnn_mean_volume


class _Subnet:
    """Packed weights of one wavelet_flow_subnetwork2D(_first) (networks.py:586-706)."""

    def __init__(self, sub, kind):
        self.normal = sub.normal
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


class _CondNet:
    """cond_network / ResidualBlock (networks.py:165-242) entirely on the tensor cores.

    The depth stencil Conv3d(1->Cm) -> PReLU -> Conv3d(Cm->1) over the (H, W, depth) volume
    (networks.py:221-225,239) is executed as two ordinary 3x3 2-D convolutions whose channel axis
    carries (depth, hidden-channel) and whose weights are the depth-banded expansion of the 3x3x3
    kernels: W1[(d,c), d'] = w1[c, :, :, d'-d+1], W2[d, (d',c)] = w2[c, :, :, d'-d+1] for |d'-d| <= 1.
    Zero padding in depth falls out of the band; zero padding in H, W is the conv's own padding."""

    def __init__(self, net, kind):
        rb = net.subnetworks[0]
        self.rb = rb
        self.c1 = tc.PackedConv(rb.conv1[0].weight, rb.conv1[0].bias, kind)
        self.ds = tc.PackedConv(rb.downsample[0].weight, rb.downsample[0].bias, kind)
        self.c2 = tc.PackedConv(rb.conv2[0].weight, rb.conv2[0].bias, kind)
        w1, b1 = rb.conv3d[0].weight.detach().float(), rb.conv3d[0].bias.detach().float()      # (Cm,1,3,3,3), (Cm)
        w2, b2 = rb.conv3d[3].weight.detach().float(), rb.conv3d[3].bias.detach().float()      # (1,Cm,3,3,3), (1)
        Cm, D = w1.shape[0], rb.out_channels
        dev = w1.device
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
        self.D = D

    def __call__(self, v8: tc.C8) -> tc.C8:
        """views (C8) -> LF condition (C8)."""
        rb = self.rb
        out = tc.conv_tc(v8, self.c1, act=ops.ACT_PRELU, slope=rb.conv1[1].weight)
        res = tc.conv_tc(v8, self.ds)
        out = tc.conv_tc(out, self.c2, act=ops.ACT_PRELU, slope=rb.relu.weight, res=res, res_mode=1)
        hid = tc.conv_tc(out, self.s1, act=ops.ACT_PRELU, slope=rb.conv3d[1].weight)
        return tc.col2im3x3_c8(tc.conv_tc(hid, self.s2g, mb=self.s2_mb), self.s2_bias, self.D)


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

    def _block(self, x, blk, training, pool):
        (p0, a0, n0), (p1, a1, n1) = blk
        x = tc.conv_tc(x, p0, act=ops.ACT_PRELU, slope=a0.weight)
        x = tc.batchnorm_c8(x, n0.weight, n0.bias, n0.running_mean, n0.running_var, batch_stats=training, eps=n0.eps)
        x = tc.conv_tc(x, p1, act=ops.ACT_PRELU, slope=a1.weight)
        return tc.batchnorm_c8(x, n1.weight, n1.bias, n1.running_mean, n1.running_var, batch_stats=training, eps=n1.eps,
                               pool=pool)

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


class CWFAEngine:
    """``engine = CWFAEngine(model); vol = engine.reconstruct(views, mean_vols)``.

    The model's parameters are packed at construction; call ``refresh()`` after changing them."""

    def __init__(self, model: CWFAModel, kind: str = "bf16"):
        if kind not in ("bf16", "fp16"):
            raise ValueError(kind)
        if not next(model.parameters()).is_cuda:
            raise RuntimeError("CWFAEngine needs the model on a CUDA device (no CPU fallback)")
        self.model, self.kind = model, kind
        self.parallel_branches = True      # LRNN and per-level coefficient nets as parallel CUDA-graph branches
        self.refresh()

    def refresh(self):
        m, kind = self.model, self.kind
        self.levels = []
        for n in range(m.n_levels):
            inn = m.conv_inn[n]
            nodes = []
            for mod in inn.module_list:
                if isinstance(mod, Fm.ConditionalAffineTransform):
                    nodes.append(("cat", mod, _Subnet(mod.subnet, kind)))
                elif isinstance(mod, Fm.PermuteRandom):
                    nodes.append(("perm", mod, 1))
                elif isinstance(mod, Fm.PermuteDim):
                    nodes.append(("perm", mod, mod.axis))
                elif isinstance(mod, (Fm.HaarTransform1D, Fm.Split)):
                    continue
                else:
                    raise NotImplementedError(f"CWFAEngine supports the default CAT graph; found {type(mod).__name__} "
                                              "(use CWFAModel.reconstruct for other block types)")
            subs = [x[2] for x in nodes if x[0] == "cat"]
            batched = None
            if all(sn.inp.Cout_p == 64 and sn.inp.Cin_p == subs[0].inp.Cin_p for sn in subs):
                # one 1x1 conv for the input layers of all sub-networks of the level (they all read the LF condition)
                ws = [(mod.subnet.block12 if mod.subnet.normal else mod.subnet.block1) for k_, mod, _ in nodes if k_ == "cat"]
                batched = tc.PackedConv(torch.cat([w.weight.detach() for w in ws], 0), torch.cat([w.bias.detach() for w in ws], 0),
                                        kind, bn=64)
            self.levels.append(dict(nodes=nodes, cond=_CondNet(m.cond_nets[n], kind), batched_in=batched))
        self.lrnn = _LRNN(m.cond_nets[-1], kind)
        self._graphs: Dict = {}

    # -------------------------------------------------------------------------------------
    def _trunks(self, n: int, v8: tc.C8):
        """Conditioning net + the trunks of all sub-networks of level n.  They depend on the conditions only (CAT
        blocks, coupling_layers.py:475-500), so forward and inverse share them and levels are independent."""
        lv = self.levels[n]
        lf8 = lv["cond"](v8)
        b_all = tc.conv_tc(lf8, lv["batched_in"]) if lv["batched_in"] is not None else None
        out, k = [], 0
        for kind, mod, extra in lv["nodes"]:
            if kind == "cat":
                out.append(("cat", mod, extra, extra.trunk(lf8, b_all, 8 * k) if b_all is not None else extra.trunk(lf8)))
                k += 1
            else:
                out.append(("perm", mod, extra, None))
        return out

    @staticmethod
    def _jac_and_tickets(n_samples: int, device, n_tickets: int = 8):
        """One zero-filled allocation (one memset) per level: the (B,) log-det accumulator plus the int32 tickets of the
        level's coupling kernels (in-kernel last-CTA finalize, cwfa_coupling_tc)."""
        buf = torch.zeros(n_samples + n_tickets, device=device, dtype=torch.float32)
        return buf[:n_samples], buf[n_samples:].view(torch.int32)

    def _couple(self, item, x, mean_vol, pending, inverse, logdet, sumsq=None, ticket=None):
        """Final conv of one sub-network with the coupling fused in its epilogue."""
        _, mod, sub, b8 = item
        perm, axis = (None, 0) if pending is None else (ops.perm_i32(pending[0], b8.data.device), pending[1])
        first = not sub.normal
        return tc.conv_tc_coupling(b8, sub.out, x, ch=mod.channels, inverse=inverse, clamp=mod.clamp,
                                   t_ext=mean_vol if first else None, t_scale=(-1.0 / math.sqrt(2)) if first else 1.0,
                                   perm=perm, perm_axis=axis, logdet=logdet, sumsq=sumsq, ticket=ticket)

    def _level_detail_inverse(self, n, v8, mean_vol, z=None):
        """Detail half ``hi`` of level n in the inverse direction and its log-det.  Independent of the other levels.
        ``z`` None = zeros (INN_z_temperature = 0, CWFA.py:906-907: z is never materialised); otherwise the latent
        sample (B, ch, H, W) of this level (``sample_z_truncated``, CWFA.py:47-64)."""
        items = self._trunks(n, v8)
        hi, pending = z, None
        jac, tickets = self._jac_and_tickets(v8.N, v8.data.device)
        k = 0
        for item in reversed(items):
            if item[0] == "cat":
                hi = self._couple(item, hi, mean_vol, pending, True, jac, ticket=tickets[k:k + 1])
                pending, k = None, (k + 1) % tickets.numel()
            elif hi is not None:
                pending = (item[1].perm_inv, item[2])       # gathered by the next coupling's epilogue
        if pending is not None:
            hi = ops.permute(hi, pending[0], pending[1])
        return hi, jac

    @torch.no_grad()
    def reconstruct(self, views: torch.Tensor, mean_vols: Sequence[Optional[torch.Tensor]], return_all: bool = False,
                    _side_streams=None, zs: Optional[Sequence[Optional[torch.Tensor]]] = None):
        """Inverse reconstruction (CWFA.py:865-924); z = 0 unless ``zs[n]`` gives level n's latent sample.

        The LRNN and the coupling coefficients of every level depend only on the views / mean volumes, not on
        each other, so under CUDA-graph capture they are issued on side streams (``_side_streams``) and become
        parallel branches of the graph; only the four Haar merges are sequential."""
        L1 = self.model.n_levels
        v8 = tc.to_c8(views, self.kind)
        mv_last = lrnn_mean_volume(mean_vols, L1)            # CWFA.py:882: mean_vols_cache[L-2] unless given explicitly
        jobs = [lambda: self.lrnn(v8, mv_last)] + [(lambda n=n: self._level_detail_inverse(n, v8, mean_vols[n], None if zs is None else zs[n])) for n in range(L1 - 1, -1, -1)]
        if _side_streams:
            main = torch.cuda.current_stream()
            fork = torch.cuda.Event()
            fork.record(main)
            results, joins = [], []
            for st, job in zip(_side_streams, jobs):
                st.wait_event(fork)
                with torch.cuda.stream(st):
                    results.append(job())
                    ev = torch.cuda.Event()
                    ev.record(st)
                joins.append(ev)
            for ev in joins:
                main.wait_event(ev)
        else:
            results = [job() for job in jobs]
        vol = results[0]
        outs, jacs = {L1: vol}, {}
        for k, n in enumerate(range(L1 - 1, -1, -1)):
            hi, jac = results[1 + k]
            vol = ops.haar1d_merge(vol, hi)          # Split^-1 + IDWT: the only sequential part of the pyramid
            outs[n], jacs[n] = vol, jac
        return (outs, jacs) if return_all else vol

    @torch.no_grad()
    def forward_nll(self, volume: torch.Tensor, views: torch.Tensor, mean_vols: Sequence[torch.Tensor]):
        """Forward pyramid + per-level NLL (CWFA.py:966-978); same outputs as CWFAModel.forward_nll."""
        v8 = tc.to_c8(views, self.kind)
        res = []
        x = volume
        for n in range(self.model.n_levels):
            lo, hi = ops.haar1d_split(x)
            jac, tickets = self._jac_and_tickets(x.shape[0], x.device)
            sumsq = torch.empty_like(jac)
            pending, k = None, 0
            for item in self._trunks(n, v8):
                if item[0] == "cat":
                    hi = self._couple(item, hi, mean_vols[n], pending, False, jac, sumsq, ticket=tickets[k:k + 1])   # sumsq: last coupling wins
                    pending, k = None, (k + 1) % tickets.numel()
                else:
                    pending = (item[1].perm, item[2])
            if pending is not None:          # trailing permutation: does not change ||z||^2
                hi = ops.permute(hi, pending[0], pending[1])
            per = (0.5 * sumsq - jac) / hi[0].numel()
            ref = (0.5 * sumsq.sum() - jac) / lo.numel()
            res.append(dict(z=hi, lo=lo, logdet=jac, sumsq=sumsq, nll_per_sample=per, nll_ref=ref))
            x = lo
        return res

    # ---- CUDA-graph replay of the whole frame -------------------------------------------
    def _graph_slot(self, views: torch.Tensor, mean_vols, slot: int = 0):
        """(graph, static_views, static_mean_vols, static_out) for this shape; ``slot`` selects an independent
        instance (own static buffers) so that several frames can be in flight."""
        key = (tuple(views.shape), tuple(None if m is None else tuple(m.shape) for m in mean_vols), views.device.index, slot)
        g = self._graphs.get(key)
        if g is None:
            sv = views.clone()
            sm = [None if m is None else m.clone() for m in mean_vols]
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    self.reconstruct(sv, sm)
            torch.cuda.current_stream().wait_stream(s)
            graph = torch.cuda.CUDAGraph()
            side = [torch.cuda.Stream() for _ in range(self.model.n_levels + 1)] if self.parallel_branches else None
            with torch.cuda.graph(graph):
                out = self.reconstruct(sv, sm, _side_streams=side)
            g = self._graphs[key] = (graph, sv, sm, out)
        return g

    def reconstruct_graphed(self, views: torch.Tensor, mean_vols: Sequence[Optional[torch.Tensor]]) -> torch.Tensor:
        """Same as ``reconstruct`` but replays a captured CUDA graph (static shapes; inputs are copied into
        static buffers, the returned tensor is the graph's static output buffer)."""
        graph, sv, sm, out = self._graph_slot(views, mean_vols)
        sv.copy_(views, non_blocking=True)
        for d, s_ in zip(sm, mean_vols):
            if d is not None:
                d.copy_(s_, non_blocking=True)
        graph.replay()
        return out

    # ---- reference-facing call with HOST buffers ------------------------------------------
    def reconstruct_host(self, views_host: torch.Tensor, mean_vols: Sequence[Optional[torch.Tensor]],
                         out_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """views on the HOST (pinned for full speed) -> H2D copy -> graph replay -> D2H copy of the
        reconstructed volume into ``out_host``.  The mean-volume conditions are dataset constants and
        stay resident on the device.  Returns ``out_host`` (synchronised)."""
        dev = mean_vols[0].device
        vd = getattr(self, "_views_dev", None)
        if vd is None or vd.shape != views_host.shape:
            vd = self._views_dev = torch.empty(views_host.shape, device=dev, dtype=torch.float32)
        vd.copy_(views_host, non_blocking=True)
        out = self.reconstruct_graphed(vd, mean_vols)
        if out_host is None:
            out_host = torch.empty(out.shape, dtype=torch.float32, pin_memory=True)
        out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_host


class StreamingReconstructor:
    """Streaming reconstruction of a sequence of frames (BASELINE.json configs[4]): ``depth`` independent graph
    instances, each on its own stream, so that (a) the H2D copy of frame i+1 and the D2H copy of frame i-1 overlap
    the replay of frame i, and (b) consecutive frames overlap on the GPU itself (the tail / dependency bubbles of
    one frame's graph are filled by the next frame's kernels).  Frames are independent, results are bit-identical
    to one-by-one reconstruction.  Inputs / outputs may live on the host (pinned) or on the device.
    The mean-volume pyramid is a dataset constant and stays on the device."""

    def __init__(self, engine: CWFAEngine, views_shape, mean_vols: Sequence[Optional[torch.Tensor]], depth: int = 2):
        self.eng, self.depth = engine, depth
        dev = mean_vols[0].device
        probe = torch.zeros(views_shape, device=dev, dtype=torch.float32)
        self.slots = [engine._graph_slot(probe, mean_vols, slot=k) for k in range(depth)]
        for (graph, sv, sm, out) in self.slots:
            for d, s_ in zip(sm, mean_vols):
                if d is not None:
                    d.copy_(s_)
        self.s_in = [torch.cuda.Stream(dev) for _ in range(depth)]
        self.s_run = [torch.cuda.Stream(dev) for _ in range(depth)]
        self.s_out = [torch.cuda.Stream(dev) for _ in range(depth)]
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_run = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        torch.cuda.synchronize(dev)

    def _staging(self, like_in: torch.Tensor, like_out: torch.Tensor):
        """Device staging buffers for host-resident frames: the H2D copy of frame i+1 and the D2H copy of frame i-1 use
        them, so a graph slot's static input/output tensors are tied up only for a device-to-device copy."""
        if getattr(self, "_stage", None) is None:
            n = self.depth + 1
            dev = like_out.device
            self._stage = ([torch.empty_like(like_in, device=dev) for _ in range(n)],
                           [torch.empty_like(like_out, device=dev) for _ in range(n)],
                           [torch.cuda.Event() for _ in range(n)], [torch.cuda.Event() for _ in range(n)],
                           [torch.cuda.Event() for _ in range(n)], [torch.cuda.Event() for _ in range(n)])
        return self._stage

    def run(self, views: Sequence[torch.Tensor], outs: Sequence[torch.Tensor], latency_events: Optional[list] = None) -> None:
        """Reconstructs ``views[i]`` into ``outs[i]``; returns when every output has been written.
        ``latency_events``: a list that receives one (start, end) pair of timing events per frame -- start when the frame's
        input copy is issued to the device queue position it can run at, end when its output has been written."""
        n = len(views)
        lat = latency_events
        cur = torch.cuda.current_stream()
        for st in self.s_in + self.s_run + self.s_out:
            st.wait_stream(cur)
        host_io = n > 0 and (not views[0].is_cuda or not outs[0].is_cuda)
        if host_io:
            _, sv0, _, out0 = self.slots[0]
            st_in, st_out, ev_h2d, ev_in_free, ev_staged, ev_out_free = self._staging(sv0, out0)
            ns = len(st_in)
            s_in, s_out = self.s_in[0], self.s_out[0]
            for i in range(n):
                k, j = i % self.depth, i % ns
                graph, sv, sm, out = self.slots[k]
                with torch.cuda.stream(s_in):
                    if i >= ns:
                        s_in.wait_event(ev_in_free[j])           # the slot that used this staging buffer has copied it in
                    if lat is not None:
                        e0 = torch.cuda.Event(enable_timing=True)
                        e0.record(s_in)
                    st_in[j].copy_(views[i], non_blocking=True)
                    ev_h2d[j].record(s_in)
                with torch.cuda.stream(self.s_run[k]):           # stream order serialises the replays of one slot
                    self.s_run[k].wait_event(ev_h2d[j])
                    sv.copy_(st_in[j], non_blocking=True)
                    ev_in_free[j].record(self.s_run[k])
                    graph.replay()
                    if i >= ns:
                        self.s_run[k].wait_event(ev_out_free[j]) # the D2H copy that used this staging buffer has finished
                    st_out[j].copy_(out, non_blocking=True)
                    ev_staged[j].record(self.s_run[k])
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_staged[j])
                    outs[i].copy_(st_out[j], non_blocking=True)
                    ev_out_free[j].record(s_out)
                    if lat is not None:
                        e1 = torch.cuda.Event(enable_timing=True)
                        e1.record(s_out)
                        lat.append((e0, e1))
        else:
            for i in range(n):
                k = i % self.depth
                graph, sv, sm, out = self.slots[k]
                with torch.cuda.stream(self.s_in[k]):
                    if i >= self.depth:
                        self.s_in[k].wait_event(self.ev_run[k])      # previous replay of this slot has consumed its input
                    if lat is not None:
                        e0 = torch.cuda.Event(enable_timing=True)
                        e0.record(self.s_in[k])
                    sv.copy_(views[i], non_blocking=True)
                    self.ev_in[k].record(self.s_in[k])
                with torch.cuda.stream(self.s_run[k]):
                    self.s_run[k].wait_event(self.ev_in[k])
                    if i >= self.depth:
                        self.s_run[k].wait_event(self.ev_out[k])     # previous output of this slot has been copied out
                    graph.replay()
                    self.ev_run[k].record(self.s_run[k])
                with torch.cuda.stream(self.s_out[k]):
                    self.s_out[k].wait_event(self.ev_run[k])
                    outs[i].copy_(out, non_blocking=True)
                    self.ev_out[k].record(self.s_out[k])
                    if lat is not None:
                        e1 = torch.cuda.Event(enable_timing=True)
                        e1.record(self.s_out[k])
                        lat.append((e0, e1))
        for st in self.s_out + self.s_run:
            cur.wait_stream(st)
        cur.synchronize()

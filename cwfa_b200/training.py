"""Training step of one flow level (BASELINE.json configs[3]).

Reference: the fine-tune branch of ``run_CWFA`` (CWFA.py:586-613 optimiser setup, :928-1015 loss / backward / step):
for flow level ``n`` the loss is ``w * MSE(gt_n, inn([z, vol_{n+1}], c, rev=True)) + (1 - w) * NLL`` with
``NLL = (0.5 * ||Z||^2 - logdet.mean()) / numel`` from the forward pass ``inn(gt_n, c)``, ``w = INN_cond_weight = 0.40984``
(main.py:107); the flow parameters and the level's conditioning net each get a Lion optimiser (lion_pytorch 0.0.7).

Every arithmetic step is one of this repo's CUDA kernels: forward and adjoint kernels through ``cwfa_b200.autograd``
(torch's autograd engine only orders them), the Lion update as ONE launch per parameter group on a flat buffer.
Data parallelism: frames are sharded over ranks, gradients are summed with ONE all-reduce of the
flat gradient buffer per optimiser (<= 4.6 MB for a flow level: a single bucket, sized for launch latency) over
``torch.distributed`` (NCCL on NVLink; gloo in the CPU tests) and the 1/world factor is folded into the Lion kernel.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import torch

from . import _lib, ops

INN_COND_WEIGHT = 0.40984      # main.py:107
_ALIGN = 4                     # floats: every parameter starts on a 16-byte boundary of the flat buffer


class FlatGroup:
    """Parameters of one optimiser group re-homed into ONE flat fp32 buffer (``p.data`` and ``p.grad`` become views), so
    that the optimiser update is one kernel launch and the data-parallel gradient reduction is one collective."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        seen, mine, loose = set(), [], []
        for p in params:
            if id(p) in seen or not p.requires_grad:
                continue
            seen.add(id(p))
            if not p.dtype.is_floating_point:
                continue
            (loose if getattr(p, "_cwfa_flat", False) else mine).append(p)
        self.params, self.loose = mine, loose          # loose: already owned by another group (e.g. the shared PReLU, networks.py:209)
        if not mine and not loose:
            raise ValueError("FlatGroup: no trainable floating-point parameters")
        dev = (mine or loose)[0].device
        offs, n = [], 0
        for p in mine:
            offs.append(n)
            n += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.offsets = offs
        self.used = [False] * len(mine)                 # "received a gradient since zero_grad": lion_pytorch skips the others
        self._mask, self._mask_key, self._hooks = None, None, []
        self._buckets, self._ov_hooks = None, []
        for i, (p, o) in enumerate(zip(mine, offs)):
            v = self.flat[o:o + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
            p.grad = self.grad[o:o + p.numel()].view(p.shape)
            p._cwfa_flat = True
            self._hooks.append(p.register_hook(lambda g, i=i: self._mark(i)))

    def _mark(self, i):
        self.used[i] = True

    # ---- gradient all-reduce overlapped with backward ("bucketed, overlapped with backward") ----------
    def enable_overlap(self, group=None, bucket_bytes: int = 32 << 20):
        """Splits the flat gradient buffer into contiguous buckets (built from the END of the buffer: backward produces the
        gradients of the last layers first) and all-reduces each bucket asynchronously the moment the last of its parameters
        has accumulated its gradient, so the collective of the early buckets runs under the rest of the backward pass
        (the LRNN step reduces 255 MB; a flow level's 4.8 MB is a single bucket).  ``finish_overlap`` waits for the buckets in
        flight and reduces whatever did not complete (parameters that received no gradient)."""
        self._ov_group = group
        self._buckets = []                     # [lo, hi, [param indices]]
        lo_hi, members, size = None, [], 0
        for i in range(len(self.params) - 1, -1, -1):
            p, o = self.params[i], self.offsets[i]
            end = o + (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
            lo_hi = (o, end) if lo_hi is None else (o, lo_hi[1])
            members.append(i)
            size += (end - o) * 4
            if size >= bucket_bytes or i == 0:
                self._buckets.append([lo_hi[0], lo_hi[1], members])
                lo_hi, members, size = None, [], 0
        self._bucket_of = {}
        for b, (_, _, mem) in enumerate(self._buckets):
            for i in mem:
                self._bucket_of[i] = b
        self._ov_hooks = [p.register_post_accumulate_grad_hook(lambda _p, i=i: self._param_ready(i)) for i, p in enumerate(self.params)]
        self._reset_overlap()

    def _reset_overlap(self):
        if getattr(self, "_buckets", None) is None:
            return
        self._pending = [len(mem) for _, _, mem in self._buckets]
        self._works = [None] * len(self._buckets)
        self._armed = False

    def arm_overlap(self):
        """The hooks launch collectives only for the backward pass that follows this call (a trainer's ``step``): a backward run
        outside a data-parallel step (gradient accumulation on one rank, a single-process comparison leg) must not issue a
        collective the other ranks never join."""
        if getattr(self, "_buckets", None) is not None:
            self._reset_overlap()
            self._armed = True

    def _param_ready(self, i):
        import torch.distributed as dist
        if not self._armed:
            return
        b = self._bucket_of[i]
        self._pending[b] -= 1
        if self._pending[b] == 0 and dist.is_available() and dist.is_initialized() and dist.get_world_size(self._ov_group) > 1:
            p, o = self.params[i], self.offsets[i]
            if p.grad is not None and p.grad.data_ptr() == self.grad.data_ptr() + 4 * o:      # gradients live in the flat buffer
                lo, hi, _ = self._buckets[b]
                self._works[b] = dist.all_reduce(self.grad[lo:hi], op=dist.ReduceOp.SUM, group=self._ov_group, async_op=True)

    def finish_overlap(self) -> int:
        """Waits for the bucket collectives in flight, reduces the buckets that never became ready; returns the number of
        collectives of this step.  Every rank sees the same ready-set (same graph), so the sequence of collectives matches."""
        import torch.distributed as dist
        n = 0
        for b, (lo, hi, _) in enumerate(self._buckets):
            if self._works[b] is not None:
                self._works[b].wait()
            else:
                dist.all_reduce(self.grad[lo:hi], op=dist.ReduceOp.SUM, group=self._ov_group)
            n += 1
        self._reset_overlap()
        return n

    def mark_all_used(self):
        """For gradients written by hand INTO the flat views (no autograd hook fires): treat every parameter as having a gradient."""
        self.used = [True] * len(self.params)

    def mask(self):
        """Byte mask over the flat buffer (1 = parameter received a gradient this step), or None when all did."""
        if all(self.used):
            return None
        key = tuple(self.used)
        if key != self._mask_key:
            m = torch.zeros(self.flat.numel(), dtype=torch.uint8, device=self.flat.device)
            for p, o, u in zip(self.params, self.offsets, self.used):
                if u:
                    m[o:o + p.numel()] = 1
            self._mask, self._mask_key = m, key
        return self._mask

    def release(self):
        """Give the parameters up (they keep their values and stay views of this buffer) so another group may re-home them."""
        for p in self.params:
            p._cwfa_flat = False
        for h in self._hooks + getattr(self, "_ov_hooks", []):
            h.remove()
        self._hooks, self._ov_hooks, self._buckets = [], [], None

    def zero_grad(self):
        self.grad.zero_()
        self.used = [False] * len(self.params)
        self._reset_overlap()
        for p, o in zip(self.params, self.offsets):           # a foreign ``zero_grad(set_to_none=True)`` may have dropped the views
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + p.numel()].view(p.shape)
        for p in self.loose:
            p.grad = None

    def grads_alias_flat(self) -> bool:
        """Autograd accumulates in place into a pre-existing ``.grad``; if anything replaced one, copy it back."""
        ok = True
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            g = p.grad
            view = self.grad[o:o + p.numel()]
            if g is None:
                # ``model.zero_grad()`` / ``p.grad = None`` (torch's set_to_none default) detached the parameter from the flat
                # buffer: no gradient this step; re-home ``.grad`` so the next backward accumulates into the buffer again
                view.zero_()
                p.grad = view.view(p.shape)
                ok = False
            elif g.data_ptr() != view.data_ptr():
                view.copy_(g.reshape(-1))
                p.grad = view.view(p.shape)
                self.used[i] = True                            # a gradient tensor was assigned by hand (or re-allocated by autograd)
                ok = False
        return ok


class Lion:
    """``lion_pytorch.Lion(params, lr=1e-4, betas=(0.9, 0.99), weight_decay=0.0)`` (v0.0.7; CWFA.py:381,608-610) with the
    update in ``cwfa_lion_step_f32``: p *= 1 - lr*wd; p -= lr*sign(b1*m + (1-b1)*g); m = b2*m + (1-b2)*g.
    ``params`` is an iterable of parameters or of group dicts ``{'params': ..., 'lr': ..., 'weight_decay': ...}``."""

    def __init__(self, params, lr: float = 1e-4, betas=(0.9, 0.99), weight_decay: float = 0.0):
        if lr <= 0.0:
            raise ValueError("lr must be positive")
        if not all(0.0 <= b <= 1.0 for b in betas):
            raise ValueError("betas must lie in [0, 1]")
        params = list(params)
        groups = params if params and isinstance(params[0], dict) else [{"params": params}]
        self.param_groups: List[Dict] = []
        for g in groups:
            fg = FlatGroup(list(g["params"]))
            self.param_groups.append(dict(params=fg.params, flat=fg, lr=float(g.get("lr", lr)), betas=tuple(g.get("betas", betas)),
                                          weight_decay=float(g.get("weight_decay", weight_decay)),
                                          exp_avg=torch.zeros_like(fg.flat), loose_state={}))
        self.grad_scale = 1.0

    def zero_grad(self, set_to_none: bool = False):
        for g in self.param_groups:
            g["flat"].zero_grad()

    def release(self):
        for g in self.param_groups:
            g["flat"].release()

    def mark_all_used(self):
        for g in self.param_groups:
            g["flat"].mark_all_used()

    def enable_overlap(self, group=None, bucket_bytes: int = 32 << 20):
        """Bucketed gradient all-reduce overlapped with backward for every flat buffer of this optimiser (``FlatGroup.enable_overlap``)."""
        for g in self.param_groups:
            g["flat"].enable_overlap(group, bucket_bytes)

    def arm_overlap(self):
        for g in self.param_groups:
            g["flat"].arm_overlap()

    def flat_grads(self) -> List[torch.Tensor]:
        for g in self.param_groups:
            g["flat"].grads_alias_flat()
        return [g["flat"].grad for g in self.param_groups]

    @torch.no_grad()
    def step(self):
        from . import packed
        packed.weights_changed()          # the kernel below rewrites the flat buffer in place: cached tensor-core executors are stale
        st = torch.cuda.current_stream().cuda_stream
        for g in self.param_groups:
            fg: FlatGroup = g["flat"]
            if not fg.flat.is_cuda:
                raise RuntimeError("cwfa_b200.Lion: parameters must live on a CUDA device (no CPU fallback)")
            fg.grads_alias_flat()
            b1, b2 = g["betas"]
            if fg.flat.numel() and any(fg.used):
                mk = fg.mask()
                _lib.call("cwfa_lion_step_masked_f32", fg.flat.data_ptr(), fg.grad.data_ptr(), g["exp_avg"].data_ptr(),
                          None if mk is None else mk.data_ptr(), fg.flat.numel(), g["lr"], b1, b2, g["weight_decay"],
                          float(self.grad_scale), st)
            for p in fg.loose:
                if p.grad is None:
                    continue
                m = g["loose_state"].setdefault(id(p), torch.zeros_like(p.data))
                gr = ops._ck(p.grad)
                _lib.call("cwfa_lion_step_f32", p.data.data_ptr(), gr.data_ptr(), m.data_ptr(), p.numel(), g["lr"], b1, b2,
                          g["weight_decay"], float(self.grad_scale), st)


class GradScaler:
    """Dynamic loss scaling with the semantics of ``torch.cuda.amp.GradScaler`` as the reference uses it
    (``GradScaler(init_scale=2.**2)``, CWFA.py:613; ``scaler.scale(loss).backward(); scaler.step(opt); scaler.update()``,
    CWFA.py:1005-1015): the loss is multiplied by ``scale`` before ``backward()``; ``step`` un-scales the gradients and SKIPS
    the optimiser step when any gradient is non-finite; ``update`` halves the scale after a skipped step and doubles it after
    ``growth_interval`` consecutive good ones.  For this package's ``Lion`` the un-scaling costs nothing (1/scale is folded into
    the update kernel's gradient factor) and the finiteness check is one sum-of-squares reduction over the flat gradient buffer."""

    def __init__(self, init_scale: float = 2.0 ** 16, growth_factor: float = 2.0, backoff_factor: float = 0.5,
                 growth_interval: int = 2000, enabled: bool = True):
        self._scale, self.growth_factor, self.backoff_factor = float(init_scale), float(growth_factor), float(backoff_factor)
        self.growth_interval, self.enabled = int(growth_interval), bool(enabled)
        self._growth_tracker, self._found_inf = 0, False
        self.skipped_steps = 0

    def get_scale(self) -> float:
        return self._scale if self.enabled else 1.0

    def scale(self, loss: torch.Tensor) -> torch.Tensor:
        return loss * self._scale if self.enabled else loss

    @staticmethod
    def _non_finite(opt: "Lion") -> bool:
        bad = False
        for g in opt.flat_grads():
            if g.numel():
                bad = bad or not bool(torch.isfinite(ops.sum_squares(g.view(1, -1))).all())     # sum g^2 is finite <=> every g is
        for pg in opt.param_groups:
            for p in pg["flat"].loose:
                if p.grad is not None:
                    bad = bad or not bool(torch.isfinite(ops.sum_squares(p.grad.reshape(1, -1))).all())
        return bad

    def step(self, optimizer: "Lion"):
        """Un-scale + step, or skip when a gradient overflowed (returns True when the step was applied)."""
        if not self.enabled:
            optimizer.step()
            return True
        if self._non_finite(optimizer):
            self._found_inf = True
            return False
        prev = optimizer.grad_scale
        optimizer.grad_scale = prev / self._scale
        try:
            optimizer.step()
        finally:
            optimizer.grad_scale = prev
        return True

    def update(self, new_scale: Optional[float] = None):
        if not self.enabled:
            return
        if new_scale is not None:
            self._scale = float(new_scale)
        elif self._found_inf:
            self._scale *= self.backoff_factor
            self._growth_tracker = 0
            self.skipped_steps += 1
        else:
            self._growth_tracker += 1
            if self._growth_tracker == self.growth_interval:
                self._scale *= self.growth_factor
                self._growth_tracker = 0
        self._found_inf = False

    def state_dict(self):
        return dict(scale=self._scale, growth_factor=self.growth_factor, backoff_factor=self.backoff_factor,
                    growth_interval=self.growth_interval, _growth_tracker=self._growth_tracker)

    def load_state_dict(self, sd):
        self._scale, self.growth_factor, self.backoff_factor = float(sd["scale"]), sd["growth_factor"], sd["backoff_factor"]
        self.growth_interval, self._growth_tracker = sd["growth_interval"], sd["_growth_tracker"]


def allreduce_gradients(optimizers: Sequence[Lion], group=None) -> int:
    """Data-parallel gradient reduction: ONE sum all-reduce per flat gradient buffer; the mean's 1/world is applied inside
    the Lion kernel (``grad_scale``) instead of in a separate pass.  Returns the number of collectives issued."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        for o in optimizers:
            o.grad_scale = 1.0
        return 0
    world = dist.get_world_size(group)
    n = 0
    for o in optimizers:
        for pg in o.param_groups:
            fg = pg["flat"]
            if getattr(fg, "_buckets", None) is not None and fg._armed:
                fg.grads_alias_flat()
                n += fg.finish_overlap()                    # buckets were reduced under the backward pass
            elif fg.grad.numel():
                fg.grads_alias_flat()
                dist.all_reduce(fg.grad, op=dist.ReduceOp.SUM, group=group)
                n += 1
        for pg in o.param_groups:
            for p in pg["flat"].loose:
                if p.grad is not None:
                    dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group)
                    n += 1
        o.grad_scale = 1.0 / world
    return n


def _maybe_overlap(optimizers, group, bucket_bytes: int = 32 << 20):
    """Data-parallel runs: reduce the gradient buckets under the backward pass."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for o in optimizers:
            o.enable_overlap(group, bucket_bytes)


def allreduce_nll_terms(sumsq: torch.Tensor, logdet: torch.Tensor, group=None):
    """Global NLL bookkeeping of a sharded batch: all-reduce of (sum ||z||^2, sum logdet, frame count) -> the three totals."""
    import torch.distributed as dist
    t = torch.stack([sumsq.sum().float(), logdet.sum().float(), torch.tensor(float(logdet.numel()), device=logdet.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, group=group)
    return t[0], t[1], t[2]


def flow_level_loss(model, n: int, gt: torch.Tensor, views: torch.Tensor, mean_vol: torch.Tensor, vol_in: torch.Tensor,
                    z: Optional[torch.Tensor] = None, cond_weight: float = INN_COND_WEIGHT):
    """Loss of flow level ``n`` exactly as CWFA.py:891-978 composes it (L2 regulariser, ``loss_func_reg='L2'``).

    gt      (B, C_n, S, S)    ground-truth volume at level n (``gt_cache[n]``)
    views   (B, 29, S, S)     normalised lenslet views
    mean_vol(B, C_n/2, S, S)  mean-volume delta condition of the level
    vol_in  (B, C_n/2, S, S)  the (detached) reconstruction of level n+1 that the inverse pass up-samples
    Returns ``(loss, parts)``; parts = dict(mse, nll, sumsq (B,), logdet (B,)).
    """
    from . import autograd as ag
    inn = model.conv_inn[n]
    cond = model.cond_nets[n](views)[-1]                                  # CWFA.py:895
    c = [cond, mean_vol]
    if z is None:
        z = torch.zeros((vol_in.shape[0],) + tuple(inn.global_out_shapes[0]), device=vol_in.device, dtype=torch.float32)
    from .modules import share_subnet_outputs
    with share_subnet_outputs():                                          # s, t of every block: computed once, used twice
        vol_rec, _ = inn([z, vol_in.detach()], c=c, rev=True)             # CWFA.py:912 (with gradients)
        mse = ag.mse_loss(gt, vol_rec)                                    # CWFA.py:953
        (Z, _lo), logdet = inn(gt, c=c)                                   # CWFA.py:966
    sumsq = ops.sum_squares(Z)
    nll = (0.5 * sumsq.sum() - logdet.mean()) / vol_rec.numel()           # CWFA.py:970,978
    loss = cond_weight * mse + (1.0 - cond_weight) * nll                  # CWFA.py:957,986
    return loss, dict(mse=mse.detach(), nll=nll.detach(), sumsq=sumsq.detach(), logdet=logdet.detach())


class _GraphedStep:
    """``graph=True`` on a trainer: the WHOLE optimisation step -- forward, backward (torch's autograd engine under stream capture),
    gradient all-reduce, Lion -- is captured once as a CUDA graph and replayed; a step is then one graph launch plus the copy of
    its inputs into static buffers (the eager step is host-bound at the full config: 14.2 ms of Python + launches per 15.9 ms
    step; the replay takes 14.3 ms).  The capture is preceded by two eager warm-up steps whose effect on the parameters, the
    optimiser state and the module buffers is undone, so the first call still performs exactly ONE step.  Re-captured when an
    input shape or a hyper-parameter (lr, betas, weight decay) changes.  Not used with a ``GradScaler`` (its skip decision is made
    on the host) nor under data parallelism (world size > 1 keeps the eager step with the bucketed all-reduce under backward).  The returned tensors are static: they are overwritten by the next step."""

    def _graph_init(self, graph: bool):
        self._graph_on, self._cg, self._cg_key, self._static, self._parts = bool(graph), None, None, None, None

    def _graph_usable(self) -> bool:
        if not self._graph_on or self.scaler is not None:
            return False
        import torch.distributed as dist        # data parallel: eager step with the bucketed all-reduce under backward (capturing the
        return not (dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1)   # NCCL calls hung on 2 GPUs)

    def _hyper_key(self):
        return tuple((g["lr"], tuple(g["betas"]), g["weight_decay"]) for o in self._optimizers() for g in o.param_groups)

    def _snapshot(self):
        snap = []
        for o in self._optimizers():
            for g in o.param_groups:
                snap += [(g["flat"].flat, g["flat"].flat.clone()), (g["exp_avg"], g["exp_avg"].clone())]
                snap += [(p.data, p.data.clone()) for p in g["flat"].loose]
                snap += [(m, m.clone()) for m in g["loose_state"].values()]
        for mod in self._trained_modules():
            snap += [(b, b.clone()) for b in mod.buffers()]
        return snap

    def _graphed(self, eager_fn, tensors):
        from . import packed
        key = (tuple(None if t is None else (tuple(t.shape), t.dtype) for t in tensors), self._hyper_key())
        if self._cg is None or key != self._cg_key:
            static = [None if t is None else t.detach().clone() for t in tensors]
            snap = self._snapshot()
            loose_before = [set(g["loose_state"]) for o in self._optimizers() for g in o.param_groups]
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(2):
                    eager_fn(*static)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            with torch.no_grad():
                for dst, src in snap:
                    dst.copy_(src)
                for ids, g in zip(loose_before, [g for o in self._optimizers() for g in o.param_groups]):
                    for k, m in g["loose_state"].items():
                        if k not in ids:
                            m.zero_()                   # momentum created during the warm-up
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                parts = eager_fn(*static)
            self._cg, self._cg_key, self._static, self._parts = graph, key, static, parts
        else:
            with torch.no_grad():
                for s_t, t in zip(self._static, tensors):
                    if t is not None and s_t.data_ptr() != t.data_ptr():
                        s_t.copy_(t)
        self._cg.replay()
        packed.weights_changed()          # the replay rewrote the parameters: packs cached by eager consumers are stale
        return self._parts


class FlowLevelTrainer(_GraphedStep):
    """One flow level's fine-tune step: Lion on the flow parameters (lr, weight decay) and Lion on the level's conditioning
    net (lr_cond), CWFA.py:596-610.  Defaults are the reference's (main.py:40-45 after the 1e-7 scaling of :238-243)."""

    def __init__(self, model, n: int, lr: float = 221e-7, lr_cond: float = 845e-7, weight_decay: float = 1e-2,
                 cond_weight: float = INN_COND_WEIGHT, group=None, precision: str = "fp32", grad_scaler: "Optional[GradScaler]" = "auto",
                 graph: bool = False):
        """``graph``: capture the whole step as a CUDA graph (``_GraphedStep``).  ``grad_scaler``: loss scaling as in the reference (``GradScaler(init_scale=4)``, CWFA.py:613,1005-1015).  ``"auto"``
        = on for ``precision='fp16'`` (the reference's autocast arithmetic: fp16 cotangents underflow without it), off for
        bf16 / fp32 (same exponent range as fp32)."""
        self.model, self.n, self.cond_weight, self.group, self.precision = model, n, cond_weight, group, precision
        self.scaler = (GradScaler(init_scale=4.0) if precision == "fp16" else None) if grad_scaler == "auto" else grad_scaler
        self.optimizer = Lion([{"params": list(model.conv_inn[n].parameters()), "lr": lr, "weight_decay": weight_decay}], lr=lr)
        self.optimizer_cond = Lion(list(model.cond_nets[n].parameters()), lr=lr_cond)
        self.collectives = 0
        self._graph_init(graph)
        if not self._graph_usable():
            _maybe_overlap([self.optimizer, self.optimizer_cond], group)   # a captured step reduces the flat buffers inside the graph

    def _optimizers(self):
        return [self.optimizer, self.optimizer_cond]

    def _trained_modules(self):
        return [self.model.conv_inn[self.n], self.model.cond_nets[self.n]]

    def release(self):
        self.optimizer.release()
        self.optimizer_cond.release()

    def step(self, gt, views, mean_vol, vol_in, z=None):
        if self._graph_usable():
            return self._graphed(self._step_eager, [gt, views, mean_vol, vol_in, z])
        return self._step_eager(gt, views, mean_vol, vol_in, z)

    def _step_eager(self, gt, views, mean_vol, vol_in, z=None):
        self.optimizer.zero_grad()
        self.optimizer_cond.zero_grad()
        self.optimizer.arm_overlap()
        self.optimizer_cond.arm_overlap()
        from . import autograd as ag
        prev = ag.set_training_precision(self.precision)      # 'bf16'/'fp16': convolutions (forward + data gradient) on tcgen05
        try:
            loss, parts = flow_level_loss(self.model, self.n, gt, views, mean_vol, vol_in, z, self.cond_weight)
            (self.scaler.scale(loss) if self.scaler is not None else loss).backward()              # CWFA.py:1007
        finally:
            ag.set_training_precision(prev)
        self.collectives = allreduce_gradients([self.optimizer, self.optimizer_cond], self.group)
        if self.scaler is not None:                                       # CWFA.py:1012-1015
            self.scaler.step(self.optimizer_cond)
            self.scaler.step(self.optimizer)
            self.scaler.update()
            parts["loss_scale"] = self.scaler.get_scale()
        else:
            self.optimizer_cond.step()                                    # CWFA.py:1002-1005
            self.optimizer.step()
        parts["loss"] = loss.detach()
        return parts


# ---------------------------------------------------------------------------------------------
# OOD decision and the coarse-to-fine fine-tune schedule
# ---------------------------------------------------------------------------------------------
def ood_decision(nll_per_level: Sequence[torch.Tensor], step_LL_to_use: int = 0, step_LL_ths_to_use: float = -1.33) -> torch.Tensor:
    """Out-of-distribution flag per frame from the forward-NLL scores (``CWFAModel.forward_nll`` -> ``nll_per_sample``).
    The reference declares the knobs (main.py:79-80: ``--step_LL_to_use`` = which flow level's likelihood, ``--step_LL_ths_to_use``
    = -1.33) but its ``evaluate_OOD_prediction`` is not in the tree (main.py:16,401 are commented out), so the rule is the
    paper's: a frame is OUT of distribution when the log-likelihood of the chosen level, LL = -NLL, falls below the threshold."""
    ll = -nll_per_level[step_LL_to_use]
    return ll < step_LL_ths_to_use


def fine_tune_flow_levels(model, frames: Sequence[dict], levels: Optional[Sequence[int]] = None, epochs_per_step: int = 1,
                          precision: str = "fp32", lr: float = 221e-7, lr_cond: float = 845e-7, weight_decay: float = 1e-2,
                          cond_weight: float = INN_COND_WEIGHT, group=None, lr_first_step: float = 80e-7):
    """Coarse-to-fine schedule of the reference's fine-tune loop (CWFA.py:746-771): the step being optimised moves from the LRNN
    (index L-1, when listed in ``levels``) through the coarsest flow (L-2) to the finest (0); while a flow level trains, its input
    volume per frame comes from the cache filled by the step below it (``upsampled_cache``, CWFA.py:748-750,919-920).

    frames: dicts with ``views`` (1,29,S,S), ``gt`` (1,D,S,S) and ``mean_vols`` (list, level n -> (1,C_n/2,S,S) = the reference's
    ``mean_vols_cache``; the LRNN gets its last entry as in CWFA.py:882 unless an extra entry -- tensor or None -- is appended).  ``levels`` defaults to the flow levels; add ``L-1`` to also run the LRNN step first.
    Returns ({step: [loss per optimiser step]}, per-frame cache of the finest reconstruction)."""
    L1 = model.n_levels
    levels = list(range(L1 - 1, -1, -1)) if levels is None else list(levels)
    with torch.no_grad():
        gt_caches = []
        for f in frames:
            _, gtc, _, _ = model.evaluate_INN_forward(f["gt"], extra_cond_in=f["mean_vols"], fix_empty_depths=False)   # GT pyramid
            gt_caches.append(gtc)
    history = {}
    from .pipeline import lrnn_mean_volume
    mv_last = lambda f: lrnn_mean_volume(f["mean_vols"], L1)          # CWFA.py:882: mean_vols_cache[L-2] unless given explicitly
    if L1 in levels:                                           # the "last step": LRNN on the coarsest ground truth
        lt = LRNNTrainer(model, lr=lr_first_step, weight_decay=weight_decay, group=group, precision=precision)
        history[L1] = []
        for _ in range(epochs_per_step):
            for i, f in enumerate(frames):
                history[L1].append(float(lt.step(gt_caches[i][L1], f["views"], mv_last(f))["loss"]))
        lt.release()
    with torch.no_grad():
        cache = [model.cond_nets[L1](f["views"], mv_last(f))[-1] for f in frames]                                        # CWFA.py:882
    for n in range(L1 - 1, -1, -1):
        if n in levels:
            tr = FlowLevelTrainer(model, n, lr=lr, lr_cond=lr_cond, weight_decay=weight_decay, cond_weight=cond_weight,
                                  group=group, precision=precision)
            history[n] = []
            for _ in range(epochs_per_step):
                for i, f in enumerate(frames):
                    history[n].append(float(tr.step(gt_caches[i][n], f["views"], f["mean_vols"][n], cache[i])["loss"]))
            tr.release()
        with torch.no_grad():                                  # last pass of the step: refill the cache for the next finer level
            for i, f in enumerate(frames):
                c0 = model.cond_nets[n](f["views"])[-1]
                z = torch.zeros((cache[i].shape[0],) + tuple(model.conv_inn[n].global_out_shapes[0]), device=cache[i].device)
                cache[i], _ = model.conv_inn[n]([z, cache[i]], c=[c0, f["mean_vols"][n]], rev=True)
    return history, cache


# ---------------------------------------------------------------------------------------------
# The LRNN ("last step") training step (CWFA.py:596-602, 882, 936-941)
# ---------------------------------------------------------------------------------------------
def lrnn_loss(model, gt: torch.Tensor, views: torch.Tensor, mean_vol: Optional[torch.Tensor] = None):
    """``F.mse_loss(curr_gt, cond_nets[-1](views, mean_vol)[-1])`` (``loss_func_first_step='L2'``).  gt: (B, D/2^(L-1), S, S).
    Differentiable through the U-Net (conv / PReLU / BatchNorm / max-pool / transposed conv adjoint kernels) and, when a mean
    volume is given, through the ConvNeXt branch and the attention gate (networks.py:548-555)."""
    from . import autograd as ag
    vol = model.cond_nets[-1](views, mean_vol)[-1]
    return ag.mse_loss(gt, vol), vol


class LRNNTrainer(_GraphedStep):
    """Lion on ``cond_nets[-1].parameters()`` with ``learning_rate_first_step`` (80e-7 after main.py:240-241) and weight decay
    1e-2 (CWFA.py:600-602); one flat buffer (<= 255 MB fp32 at the full config), one all-reduce per step under data parallelism."""

    def __init__(self, model, lr: float = 80e-7, weight_decay: float = 1e-2, group=None, precision: str = "fp32",
                 grad_scaler: "Optional[GradScaler]" = "auto", graph: bool = False):
        self.model, self.group, self.precision = model, group, precision
        self.scaler = (GradScaler(init_scale=4.0) if precision == "fp16" else None) if grad_scaler == "auto" else grad_scaler
        self.optimizer = Lion([{"params": list(model.cond_nets[-1].parameters()), "lr": lr, "weight_decay": weight_decay}], lr=lr)
        self.collectives = 0
        self._graph_init(graph)
        if not self._graph_usable():
            _maybe_overlap([self.optimizer], group)

    def _optimizers(self):
        return [self.optimizer]

    def _trained_modules(self):
        return [self.model.cond_nets[-1]]

    def release(self):
        self.optimizer.release()

    def step(self, gt, views, mean_vol=None):
        if self._graph_usable():
            return self._graphed(self._step_eager, [gt, views, mean_vol])
        return self._step_eager(gt, views, mean_vol)

    def _step_eager(self, gt, views, mean_vol=None):
        from . import autograd as ag
        self.optimizer.zero_grad()
        self.optimizer.arm_overlap()
        prev = ag.set_training_precision(self.precision)
        try:
            loss, _ = lrnn_loss(self.model, gt, views, mean_vol)
            (self.scaler.scale(loss) if self.scaler is not None else loss).backward()
        finally:
            ag.set_training_precision(prev)
        self.collectives = allreduce_gradients([self.optimizer], self.group)
        if self.scaler is not None:
            self.scaler.step(self.optimizer)
            self.scaler.update()
        else:
            self.optimizer.step()
        return dict(loss=loss.detach())

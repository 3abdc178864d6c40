"""Graph container with the FrEIA ``framework`` API surface the CWFA path uses.

Mirrors (API-compatible, independently written): ``Node``, ``InputNode``, ``ConditionNode``,
``OutputNode``, ``GraphINN`` (+ the deprecated ``ReversibleGraphNet`` alias) and
``SequenceINN`` of the reference's vendored FrEIA (FrEIA/framework/graph_inn.py:13-326,
sequence_inn.py:10-99, reversible_graph_net.py).  Semantics kept:

* ``Node(inputs, module_type, module_args, conditions=None, name=None)`` builds its module as
  ``module_type(input_shapes, dims_c=cond_shapes, **args)`` (or without ``dims_c`` when there
  are no conditions) and infers ``output_dims`` (graph_inn.py:63-74).
* ``GraphINN(nodes)(x_or_z, c=[...], rev=False, jac=True)`` executes nodes in (reverse)
  dependency order, sums per-node log-dets into a ``(B,)`` tensor, returns a single tensor
  when there is one output unless ``force_tuple_output`` (graph_inn.py:242-326).
* ``module_list`` holds the node modules in execution order so ``state_dict`` keys are
  ``module_list.{i}.…`` exactly as in reference checkpoints.
* Conditions are matched to ``ConditionNode``s in the order they appear in the node list.

The graph walk stays in Python (about 13 nodes per level); the arithmetic inside every node is
a CUDA kernel launch through the C ABI.
"""
from __future__ import annotations

import warnings
from typing import Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from .modules import InvertibleModule

NodeRef = Tuple["Node", int]


class Node:
    """One transformation of the graph with any number of inputs and outputs."""

    def __init__(self, inputs, module_type, module_args: Optional[dict] = None, conditions=None, name=None):
        self.name = name if name else hex(id(self))[-6:]
        self.inputs: List[NodeRef] = self._normalise_inputs(inputs)
        if conditions is None:
            conditions = []
        self.conditions = list(conditions) if isinstance(conditions, (list, tuple)) else [conditions]
        self.module_type = module_type
        self.module_args = dict(module_args or {})
        self.outputs: List[Optional[NodeRef]] = []

        self.input_dims = [src.output_dims[k] for src, k in self.inputs]
        self.condition_dims = [cn.output_dims[0] for cn in self.conditions]
        self.module, self.output_dims = self.build_module(self.condition_dims, self.input_dims)

        for slot, (src, k) in enumerate(self.inputs):
            src.outputs[k] = (self, slot)
        for k in range(len(self.output_dims)):
            setattr(self, f"out{k}", (self, k))
            self.outputs.append(None)

    def build_module(self, condition_shapes, input_shapes):
        if len(self.conditions) > 0:
            module = self.module_type(input_shapes, dims_c=condition_shapes, **self.module_args)
        else:
            module = self.module_type(input_shapes, **self.module_args)
        return module, module.output_dims(input_shapes)

    def _normalise_inputs(self, inputs) -> List[NodeRef]:
        if isinstance(inputs, Node):
            return [(inputs, 0)]
        if isinstance(inputs, (list, tuple)):
            if len(inputs) == 0:
                return list(inputs)
            if isinstance(inputs[0], (list, tuple)):
                return [tuple(i) for i in inputs]
            if len(inputs) == 2 and isinstance(inputs[0], Node):
                return [(inputs[0], int(inputs[1]))]
            raise RuntimeError(f"Cannot parse inputs provided to node '{self.name}'.")
        raise ValueError(f"Received object of invalid type ({type(inputs)}) as input for node '{self.name}'.")

    def __repr__(self):
        mt = self.module_type.__name__ if self.module_type is not None else ""
        return f"{type(self).__name__} {self.name!r}: {self.input_dims} -> {mt} -> {self.output_dims}"


class _SpecialNode(Node):
    def build_module(self, condition_shapes, input_shapes):
        if len(condition_shapes) > 0:
            raise ValueError(f"{type(self).__name__} does not accept conditions")
        return None, self._dims(input_shapes)


class InputNode(_SpecialNode):
    """Input of the whole net (output when run in reverse)."""

    def __init__(self, *dims: int, name=None):
        self.dims = tuple(dims)
        super().__init__([], None, {}, name=name)

    def _dims(self, input_shapes):
        return [self.dims]


class ConditionNode(_SpecialNode):
    """Conditional input routed to the sub-networks of coupling blocks."""

    def __init__(self, *dims: int, name=None):
        self.dims = tuple(dims)
        super().__init__([], None, {}, name=name)
        self.outputs = []

    def _dims(self, input_shapes):
        return [self.dims]


class OutputNode(_SpecialNode):
    """Output of the whole net (input when run in reverse)."""

    def __init__(self, in_node, name=None):
        super().__init__(in_node, None, {}, name=name)

    def _dims(self, input_shapes):
        if len(input_shapes) != 1:
            raise ValueError(f"Output node received {len(input_shapes)} inputs, but only single input is allowed.")
        return []


def topological_order(all_nodes: Sequence[Node], in_nodes: Sequence[Node], out_nodes: Sequence[Node]) -> List[Node]:
    """Dependency order that matches the reference's for the graphs CWFA builds
    (graph_inn.py:429-473: breadth-first peel from the outputs, then reversed), so that
    ``module_list`` indices -- and therefore checkpoint keys -- line up."""
    consumers = {id(n): set() for n in all_nodes}
    producers = {id(n): [] for n in all_nodes}
    by_id = {id(n): n for n in all_nodes}
    for n in all_nodes:
        for src, _ in n.inputs:
            if id(src) not in by_id:
                raise ValueError(f"{n} gets input from {src}, but the latter is not in the node list.")
            if id(src) not in [id(p) for p in producers[id(n)]]:
                producers[id(n)].append(src)
            consumers[id(src)].add(id(n))
    peeled: List[Node] = []
    frontier = list(out_nodes)
    while frontier:
        node = frontier.pop(0)
        peeled.append(node)
        for src in producers[id(node)]:
            consumers[id(src)].discard(id(node))
            if not consumers[id(src)]:
                frontier.append(src)
    for n in in_nodes:
        if not any(n is p for p in peeled):
            raise ValueError(f"Error in graph: {n} is not connected to any output.")
    if any(consumers[id(n)] for n in all_nodes):
        raise ValueError("Graph is cyclic.")
    return peeled[::-1]


class GraphINN(InvertibleModule):
    """Invertible network assembled from ``Node``s; run forward or (``rev=True``) backward."""

    def __init__(self, node_list, force_tuple_output=False, verbose=False):
        node_list = list(node_list)
        in_nodes = [n for n in node_list if isinstance(n, InputNode)]
        out_nodes = [n for n in node_list if isinstance(n, OutputNode)]
        condition_nodes = [n for n in node_list if isinstance(n, ConditionNode)]
        for n in node_list:
            for dst in n.outputs:
                if dst is not None and not any(dst[0] is m for m in node_list):
                    raise ValueError(f"{dst[0]} gets input from {n}, but it is not in the node list passed to GraphINN.")
        ordered = topological_order(node_list, in_nodes, out_nodes)
        super().__init__([n.output_dims[0] for n in in_nodes], [n.output_dims[0] for n in condition_nodes])
        self.node_list = ordered
        self.in_nodes = in_nodes
        self.out_nodes = out_nodes
        self.condition_nodes = condition_nodes
        self.global_out_shapes = [n.input_dims[0] for n in out_nodes]
        self.force_tuple_output = force_tuple_output
        self.module_list = nn.ModuleList([n.module for n in ordered if n.module is not None])
        if verbose:
            print(self)

    def output_dims(self, input_dims):
        if len(self.global_out_shapes) == 1 and not self.force_tuple_output:
            raise ValueError("You can only call output_dims on a GraphINN with more than one output "
                             "or when setting force_tuple_output=True.")
        return self.global_out_shapes

    def forward(self, x_or_z, c=None, rev: bool = False, jac: bool = True, intermediate_outputs: bool = False, x=None):
        if x is not None:
            x_or_z = x
            warnings.warn("You called GraphINN(x=...). x is now called x_or_z, please pass input as positional argument.")
        if torch.is_tensor(x_or_z):
            x_or_z = (x_or_z,)
        if torch.is_tensor(c):
            c = (c,)
        c = [] if c is None else list(c)
        starts = self.out_nodes if rev else self.in_nodes
        if len(x_or_z) != len(starts):
            raise ValueError(f"Got {len(x_or_z)} inputs, but expected {len(starts)}.")
        if len(c) != len(self.condition_nodes):
            raise ValueError(f"Got {len(c)} conditions, but expected {len(self.condition_nodes)}.")

        first = x_or_z[0]
        total_jac = torch.zeros(first.shape[0], dtype=first.dtype, device=first.device)
        values = {}
        jac_by_node = {} if jac else None
        for t, n in zip(x_or_z, starts):
            values[(id(n), 0)] = t
        for t, n in zip(c, self.condition_nodes):
            values[(id(n), 0)] = t

        for node in (reversed(self.node_list) if rev else self.node_list):
            if node.module is None:
                continue
            links = node.outputs if rev else node.inputs
            mod_in = tuple(values[(id(src), k)] for src, k in links)
            if node.conditions:
                mod_c = tuple(values[(id(cn), 0)] for cn in node.conditions)
                result = node.module(mod_in, c=mod_c, rev=rev, jac=jac)
            else:
                result = node.module(mod_in, rev=rev, jac=jac)
            outs, node_jac = self._check_output(node, result, jac, rev)
            for k, t in enumerate(outs):
                values[(id(node), k)] = t
            if jac:
                total_jac = total_jac + node_jac
                jac_by_node[node] = node_jac

        ends = self.in_nodes if rev else self.out_nodes
        for n in ends:
            src, k = (n.outputs if rev else n.inputs)[0]
            values[(id(n), 0)] = values[(id(src), k)]
        if intermediate_outputs:
            lookup = {(n, k): v for n in self.node_list + self.condition_nodes
                      for (nid, k), v in values.items() if nid == id(n)}
            return lookup, jac_by_node
        result = [values[(id(n), 0)] for n in ends]
        if len(result) == 1 and not self.force_tuple_output:
            return result[0], total_jac
        return tuple(result), total_jac

    def _check_output(self, node, result, jac, rev):
        if torch.is_tensor(result):
            raise ValueError(f"The node {node}'s module returned a tensor only; it must return (outputs, jac).")
        if len(result) != 2:
            raise ValueError(f"The node {node}'s module returned a tuple of length {len(result)}, "
                             "but should return a tuple `z_or_x, jac`.")
        outs, node_jac = result
        if torch.is_tensor(outs):
            raise ValueError(f"The node {node}'s module returns a tensor; it must return a sequence of tensors.")
        expected = len(node.inputs if rev else node.outputs)
        if len(outs) != expected:
            raise ValueError(f"The node {node}'s module returned {len(outs)} output variables, but should return {expected}.")
        if not torch.is_tensor(node_jac):
            if isinstance(node_jac, (float, int)):
                node_jac = torch.zeros(outs[0].shape[0], dtype=outs[0].dtype, device=outs[0].device) + node_jac
            elif jac:
                raise ValueError(f"The node {node}'s module returned a non-tensor as Jacobian: {node_jac}")
            elif node_jac is not None:
                raise ValueError(f"The node {node}'s module returned neither None nor a Jacobian: {node_jac}")
        return outs, node_jac

    def log_jacobian_numerical(self, x, c=None, rev=False, h=1e-04):
        """Central finite-difference log|det J| (independent log-det check, graph_inn.py:369-407)."""
        xs = list(x) if isinstance(x, (list, tuple)) else [x]
        B = xs[0].shape[0]
        sizes = [int(np.prod(t.shape[1:])) for t in xs]
        n = sum(sizes)
        flat = torch.cat([t.reshape(B, -1) for t in xs], dim=1)

        def run(v):
            parts = torch.split(v, sizes, dim=1)
            parts = [p.reshape(t.shape) for p, t in zip(parts, xs)]
            arg = parts if isinstance(x, (list, tuple)) else parts[0]
            y, _ = self.forward(arg, c=c, rev=rev, jac=False)
            ys = list(y) if isinstance(y, (list, tuple)) else [y]
            return torch.cat([t.reshape(B, -1) for t in ys], dim=1)

        J = torch.zeros(B, n, n, dtype=torch.float64)
        for i in range(n):
            d = torch.zeros_like(flat)
            d[:, i] = h
            J[:, :, i] = ((run(flat + d) - run(flat - d)) / (2 * h)).double().cpu()
        return torch.stack([torch.slogdet(J[b])[1] for b in range(B)]).to(xs[0].dtype).to(xs[0].device)

    def get_node_by_name(self, name) -> Optional[Node]:
        for n in self.node_list:
            if n.name == name:
                return n
        return None

    def get_module_by_name(self, name) -> Optional[nn.Module]:
        n = self.get_node_by_name(name)
        return None if n is None else n.module


class ReversibleGraphNet(GraphINN):
    """Deprecated alias kept for API parity (FrEIA/framework/reversible_graph_net.py:9-36)."""

    def __init__(self, node_list, ind_in=None, ind_out=None, verbose=True, force_tuple_output=False):
        warnings.warn("ReversibleGraphNet is deprecated in favour of GraphINN.", DeprecationWarning)
        if ind_in is not None or ind_out is not None:
            raise ValueError("ind_in / ind_out are not supported; pass the node list only.")
        super().__init__(node_list, verbose=verbose, force_tuple_output=force_tuple_output)


class SequenceINN(InvertibleModule):
    """Linear chain of invertible modules (FrEIA/framework/sequence_inn.py:10-99)."""

    def __init__(self, *dims: int, force_tuple_output=False):
        super().__init__([dims])
        self.shapes = [tuple(dims)]
        self.conditions = []
        self.module_list = nn.ModuleList()
        self.force_tuple_output = force_tuple_output

    def append(self, module_class, cond=None, cond_shape=None, **kwargs):
        dims_in = [self.shapes[-1]]
        self.conditions.append(cond)
        if cond is not None:
            kwargs["dims_c"] = [cond_shape]
        module = module_class(dims_in, **kwargs)
        self.module_list.append(module)
        out = module.output_dims(dims_in)
        assert len(out) == 1, "Module has more than one output"
        self.shapes.append(out[0])

    def __getitem__(self, item):
        return self.module_list[item]

    def __len__(self):
        return len(self.module_list)

    def __iter__(self):
        return iter(self.module_list)

    def output_dims(self, input_dims=None):
        if not self.force_tuple_output:
            raise ValueError("You can only call output_dims on a SequenceINN when setting force_tuple_output=True.")
        return input_dims                      # sequence_inn.py:62-66 returns its argument unchanged

    def forward(self, x_or_z, c: Iterable[torch.Tensor] = None, rev: bool = False, jac: bool = True):
        """sequence_inn.py:68-99: a tensor or a 1-tuple in; the log-det starts as the integer 0 and takes the type of the
        first module's ``jac`` (python number for fixed transforms, ``(B,)`` tensor for couplings)."""
        order = range(len(self.module_list))
        order = reversed(order) if rev else order
        log_det = 0
        cur = (x_or_z,) if torch.is_tensor(x_or_z) else x_or_z
        for i in order:
            if self.conditions[i] is None:
                cur, j = self.module_list[i](cur, jac=jac, rev=rev)
            else:
                cur, j = self.module_list[i](cur, c=[c[self.conditions[i]]], jac=jac, rev=rev)
            log_det = j + log_det
        return (cur if self.force_tuple_output else cur[0]), log_det

"""MessagePassing-compatible operator modules (the reference's ``operator_registry`` slot,
/root/reference/src/train_gnn_embeddings.py:96-102; constructed as ``operator((-1, -1), C)`` at
src/models/models_graph.py:17,23) backed by the agx kernels.

Parameter names and shapes follow PyG 2.0.2 so state-dicts interchange with the reference:
``lin_l.weight [out, F_src]``, ``lin_l.bias [out]``, ``lin_r.weight [out, F_dst]`` (SAGEConv);
``lin_rel`` / ``lin_root`` (GraphConv, with ``lin_l`` / ``lin_r`` aliases).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple, Union

import torch
import torch.nn as nn

from . import functional as AF
from .functional import ConvSpec, RelSpec
from .graph import HeteroPlan, get_plan


class Linear(nn.Module):
    """PyG-style ``Linear`` with lazy (-1) input size (src/models/models_graph.py:18)."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True,
                 weight_initializer: Optional[str] = None):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.weight_initializer = weight_initializer     # 'glorot': PyG's GATConv projections
        if in_channels > 0:
            self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        else:
            self.weight = nn.parameter.UninitializedParameter()
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    @property
    def is_lazy(self) -> bool:
        return isinstance(self.weight, nn.parameter.UninitializedParameter)

    def reset_parameters(self):
        if self.in_channels > 0 and not self.is_lazy:
            bound = 1.0 / math.sqrt(self.in_channels)
            with torch.no_grad():
                if self.weight_initializer == 'glorot':
                    g = math.sqrt(6.0 / (self.in_channels + self.out_channels))
                    self.weight.uniform_(-g, g)
                else:
                    self.weight.uniform_(-bound, bound)
                if self.bias is not None:
                    self.bias.uniform_(-bound, bound)

    def materialize(self, in_channels: int, device=None):
        if self.is_lazy:
            self.in_channels = int(in_channels)
            dev = device if device is not None else (
                self.bias.device if self.bias is not None else None)
            self.weight.materialize((self.out_channels, self.in_channels), device=dev)
            self.reset_parameters()

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        if self.is_lazy:
            destination[prefix + 'weight'] = self.weight
            if self.bias is not None:
                destination[prefix + 'bias'] = self.bias if keep_vars else self.bias.detach()
        else:
            super()._save_to_state_dict(destination, prefix, keep_vars)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                              unexpected_keys, error_msgs):
        w = state_dict.get(prefix + 'weight', None)
        if isinstance(w, nn.parameter.UninitializedParameter):
            if not self.is_lazy:
                error_msgs.append(f'{prefix}weight: cannot load a lazy weight into an initialised '
                                  f'Linear')
            if self.bias is not None and prefix + 'bias' in state_dict:
                with torch.no_grad():
                    self.bias.copy_(state_dict[prefix + 'bias'])
            return
        if w is not None and self.is_lazy:
            self.in_channels = w.shape[-1]
            dev = self.bias.device if self.bias is not None else w.device
            self.weight.materialize((self.out_channels, self.in_channels), device=dev)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys,
                                      unexpected_keys, error_msgs)

    def _apply(self, fn, recurse=True):
        # nn.Module._apply would touch the uninitialised weight; move the rest only
        if self.is_lazy:
            if self.bias is not None:
                with torch.no_grad():
                    self.bias.data = fn(self.bias.data)
            return self
        return super()._apply(fn, recurse)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self.materialize(x.shape[-1], x.device)
        return AF.fused_linear([x], self.weight, self.bias)


class MessagePassing(nn.Module):
    """Base class recognised by ``to_hetero`` (PyG tests ``isinstance(m, MessagePassing)``)."""

    aggr = 'add'

    def _lins(self) -> Tuple[Linear, Optional[Linear]]:
        raise NotImplementedError

    def reset_parameters(self):
        for lin in self._lins():
            if lin is not None:
                lin.reset_parameters()

    def rel_params(self, f_src: int, f_dst: int, device):
        """(W_l, b_l or None, W_r or None) with lazy sizes resolved."""
        lin_l, lin_r = self._lins()
        lin_l.materialize(f_src, device)
        if lin_r is not None:
            lin_r.materialize(f_dst, device)
        return lin_l.weight, lin_l.bias, (lin_r.weight if lin_r is not None else None)

    def forward(self, x: Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]],
                edge_index: torch.Tensor, size=None) -> torch.Tensor:
        """``x``: Tensor or ``(x_src, x_dst)``; ``edge_index`` int64 [2, E], row 0 = source,
        row 1 = destination.  Returns ``[N_dst, out_channels]``."""
        if torch.is_tensor(x):
            x = (x, x)
        x_src, x_dst = x
        same = x_src is x_dst
        n_src, n_dst = x_src.shape[0], x_dst.shape[0]
        nodes = {'s': n_src} if same else {'s': n_src, 'd': n_dst}
        key = ('s', 'r', 's' if same else 'd')
        plan = get_plan({key: edge_index}, nodes)
        wl, bl, wr = self.rel_params(x_src.shape[1], x_dst.shape[1], x_src.device)
        # a standalone operator call is ONE relation: the layer-level C entry points
        # (agx_sage_layer_fwd / _bwd); to_hetero fuses all relations of a layer instead
        return AF.sage_layer(plan[key], self.aggr == 'mean', x_src, x_dst, wl, bl, wr)


class SAGEConv(MessagePassing):
    """``out = lin_l(mean_{j in N(i)} x_src[j]) + lin_r(x_dst[i])`` (PyG 2.0.2 SAGEConv, SURVEY a-4)."""

    def __init__(self, in_channels, out_channels: int, normalize: bool = False,
                 root_weight: bool = True, bias: bool = True, aggr: str = 'mean'):
        super().__init__()
        if normalize:
            raise NotImplementedError('SAGEConv(normalize=True) is not used by the reference')
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        if aggr not in ('mean', 'add', 'sum'):
            raise NotImplementedError(f'aggr={aggr}')
        self.aggr = 'mean' if aggr == 'mean' else 'add'
        self.in_channels, self.out_channels = in_channels, out_channels
        self.root_weight = root_weight
        self.lin_l = Linear(in_channels[0], out_channels, bias=bias)
        if root_weight:
            self.lin_r = Linear(in_channels[1], out_channels, bias=False)

    def _lins(self):
        return self.lin_l, (self.lin_r if self.root_weight else None)


class GraphConv(MessagePassing):
    """``out = lin_rel(sum_{j in N(i)} x_src[j]) + lin_root(x_dst[i])`` (PyG 2.0.2 GraphConv, a-5)."""

    def __init__(self, in_channels, out_channels: int, aggr: str = 'add', bias: bool = True):
        super().__init__()
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        if aggr not in ('mean', 'add', 'sum'):
            raise NotImplementedError(f'aggr={aggr}')
        self.aggr = 'mean' if aggr == 'mean' else 'add'
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_rel = Linear(in_channels[0], out_channels, bias=bias)
        self.lin_root = Linear(in_channels[1], out_channels, bias=False)

    @property
    def lin_l(self):
        return self.lin_rel

    @property
    def lin_r(self):
        return self.lin_root

    def _lins(self):
        return self.lin_rel, self.lin_root


class GATConv(MessagePassing):
    """PyG 2.0.2 ``GATConv`` with the reference's settings (one head, ``negative_slope=0.2``, no
    attention dropout, self loops added): the default ``--operator`` of
    /root/reference/src/train_gnn_embeddings.py:15,99.  Parameters ``lin_l.weight``,
    ``lin_r.weight`` (no bias), ``att_l``, ``att_r`` ``[1, 1, C]``, ``bias [C]``.

        x_l = lin_l(x_src); a_l = <x_l, att_l>; a_r = <lin_r(x_dst), att_r>
        out_i = sum_j softmax_i(leaky_relu(a_l[j] + a_r[i])) x_l[j] + bias
    """

    aggr = 'add'

    def __init__(self, in_channels, out_channels: int, heads: int = 1, concat: bool = True,
                 negative_slope: float = 0.2, dropout: float = 0.0, add_self_loops: bool = True,
                 bias: bool = True):
        super().__init__()
        if heads != 1 or not concat or dropout != 0.0:
            raise NotImplementedError('GATConv: only heads=1, concat=True, dropout=0 (the '
                                      'reference constructs operator((-1, -1), C) with defaults)')
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.heads, self.negative_slope, self.add_self_loops = 1, negative_slope, add_self_loops
        self.lin_l = Linear(in_channels[0], out_channels, bias=False, weight_initializer='glorot')
        self.lin_r = Linear(in_channels[1], out_channels, bias=False, weight_initializer='glorot')
        self.att_l = nn.Parameter(torch.empty(1, 1, out_channels))
        self.att_r = nn.Parameter(torch.empty(1, 1, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def _lins(self):
        return self.lin_l, self.lin_r

    def reset_parameters(self):
        super().reset_parameters()
        nn.init.xavier_uniform_(self.att_l)
        nn.init.xavier_uniform_(self.att_r)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def gat_params(self, f_src: int, f_dst: int, device):
        """(W_l, W_r, att_l, att_r, bias or None) with lazy sizes resolved."""
        self.lin_l.materialize(f_src, device)
        self.lin_r.materialize(f_dst, device)
        return self.lin_l.weight, self.lin_r.weight, self.att_l, self.att_r, self.bias

    def forward(self, x, edge_index: torch.Tensor, size=None) -> torch.Tensor:
        if torch.is_tensor(x):
            x = (x, x)
        x_src, x_dst = x
        same = x_src is x_dst
        plan = AF.GATPlan.get(edge_index, x_src.shape[0], x_dst.shape[0], self.add_self_loops)
        params = [p for p in self.gat_params(x_src.shape[1], x_dst.shape[1], x_src.device)
                  if p is not None]
        from .hetero import _is_identity_input          # one-hot node features: I W^T = W^T
        types = ['s'] if same else ['s', 'd']
        xs = [x_src.contiguous()] if same else [x_src.contiguous(), x_dst.contiguous()]
        spec = AF.GATSpec(
            node_types=types,
            rels=[AF.GATRelSpec(plan, 's', 's' if same else 'd', 0, 1, 2, 3,
                                4 if self.bias is not None else -1)],
            out_channels=self.out_channels, slope=float(self.negative_slope),
            identity={t: (not xt.requires_grad) and _is_identity_input(xt)
                      for t, xt in zip(types, xs)})
        (out,) = AF.hetero_gat(spec, xs, params)
        return out

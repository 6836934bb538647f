"""CPU restatement of the reference's hetero message-passing path.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py; "parity unpinned" for the PyG-owned arithmetic).

Every function follows a reference call site (paths relative to /root/reference) and, where the
arithmetic lives in the absent third-party PyG 2.0.2 / torch-scatter 2.0.9, PyG's published
algorithm on the same ATen CPU kernels the reference dispatches to.
"""
from __future__ import annotations

import copy
import math
from collections import OrderedDict, defaultdict, deque
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def key2str(key) -> str:
    return '__'.join(key) if isinstance(key, tuple) else key


# --------------------------------------------------------------------------------------------
# a-2  T.ToUndirected()  -- call site src/train_gnn_embeddings.py:117-120
# --------------------------------------------------------------------------------------------
def to_undirected(edge_index_dict) -> "OrderedDict":
    """PyG 2.0.2 ``ToUndirected`` on a heterograph.  Bipartite stores (src != dst) get a new
    store ``(dst, 'rev_'+rel, src)`` holding ``stack([col, row])`` with the edge order kept;
    a non-bipartite store (teacher_rel) is replaced in place by
    ``coalesce(cat([row, col]), cat([col, row]))``: unique pairs in ascending ``row*N+col`` order.
    New stores are appended after all originals (insertion order of the store dict)."""
    out = OrderedDict()
    rev = OrderedDict()
    for (src, rel, dst), ei in edge_index_dict.items():
        row, col = ei[0], ei[1]
        if src != dst:
            out[(src, rel, dst)] = ei
            rev[(dst, 'rev_' + rel, src)] = torch.stack([col, row], dim=0)
        else:
            r2 = torch.cat([row, col])
            c2 = torch.cat([col, row])
            n = int(max(int(r2.max()), int(c2.max()))) + 1 if r2.numel() else 1
            idx = torch.unique(r2 * n + c2, sorted=True)
            out[(src, rel, dst)] = torch.stack([idx // n, idx % n], dim=0)
    out.update(rev)
    return out


# --------------------------------------------------------------------------------------------
# K1 oracle: CSR (by destination) / CSC (by source) with the neighbour order of the edge list
# --------------------------------------------------------------------------------------------
def csr_build(keys: np.ndarray, vals: np.ndarray, n_rows: int):
    """Stable counting sort of the edge list by ``keys``.  Returns (rowptr[int64 n_rows+1],
    col = vals[perm], eid = perm).  Within a row, neighbours keep edge-list order, which is the
    order CPU ``scatter_add_`` accumulates in."""
    keys = np.asarray(keys, dtype=np.int64)
    vals = np.asarray(vals, dtype=np.int64)
    perm = np.argsort(keys, kind='stable')
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(np.bincount(keys, minlength=n_rows), out=rowptr[1:])
    return rowptr, vals[perm], perm


# --------------------------------------------------------------------------------------------
# a-6  MessagePassing.propagate for a dense edge_index: gather + torch_scatter.scatter
# --------------------------------------------------------------------------------------------
def propagate(x_src: torch.Tensor, edge_index: torch.Tensor, n_dst: int, reduce: str):
    x_j = x_src.index_select(0, edge_index[0])
    idx = edge_index[1].view(-1, 1).expand_as(x_j)
    out = torch.zeros(n_dst, x_src.shape[1], dtype=x_src.dtype).scatter_add_(0, idx, x_j)
    if reduce == 'mean':
        ones = torch.ones(edge_index.shape[1], dtype=x_src.dtype)
        count = torch.zeros(n_dst, dtype=x_src.dtype).scatter_add_(0, edge_index[1], ones)
        count[count < 1] = 1
        out = out / count.view(-1, 1)
    elif reduce not in ('add', 'sum'):
        raise ValueError(reduce)
    return out


# --------------------------------------------------------------------------------------------
# PyG ``Linear`` with lazy (-1) input size -- used at src/models/models_graph.py:18 and inside convs
# --------------------------------------------------------------------------------------------
class Linear(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, bias: bool = True,
                 weight_initializer: Optional[str] = None):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.weight_initializer = weight_initializer
        if in_channels > 0:
            self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        else:
            self.weight = nn.parameter.UninitializedParameter()
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.in_channels > 0:
            bound = 1.0 / math.sqrt(self.in_channels)
            with torch.no_grad():
                if self.weight_initializer == 'glorot':  # PyG inits.glorot: U(+-sqrt(6/(in+out)))
                    g = math.sqrt(6.0 / (self.in_channels + self.out_channels))
                    self.weight.uniform_(-g, g)
                else:
                    self.weight.uniform_(-bound, bound)      # kaiming_uniform(a=sqrt(5))
                if self.bias is not None:
                    self.bias.uniform_(-bound, bound)

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        if isinstance(self.weight, nn.parameter.UninitializedParameter):
            destination[prefix + 'weight'] = self.weight      # PyG's _lazy_save_hook behaviour
            if self.bias is not None:
                destination[prefix + 'bias'] = self.bias if keep_vars else self.bias.detach()
        else:
            super()._save_to_state_dict(destination, prefix, keep_vars)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                              unexpected_keys, error_msgs):
        w = state_dict.get(prefix + 'weight', None)
        lazy_here = isinstance(self.weight, nn.parameter.UninitializedParameter)
        if isinstance(w, nn.parameter.UninitializedParameter):
            if not lazy_here:
                error_msgs.append(f'{prefix}weight: cannot load a lazy weight into an '
                                  f'initialised Linear')
            if self.bias is not None and prefix + 'bias' in state_dict:
                with torch.no_grad():
                    self.bias.copy_(state_dict[prefix + 'bias'])
            return
        if w is not None and lazy_here:
            self.in_channels = w.shape[-1]
            self.weight.materialize((self.out_channels, self.in_channels))
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys,
                                      unexpected_keys, error_msgs)

    def forward(self, x):
        if isinstance(self.weight, nn.parameter.UninitializedParameter):
            self.in_channels = x.shape[-1]
            self.weight.materialize((self.out_channels, self.in_channels))
            self.reset_parameters()
        return F.linear(x, self.weight, self.bias)


class MessagePassing(nn.Module):
    """Marker base class: ``to_hetero`` duplicates instances per edge type."""


# a-4  SAGEConv -- operator slot src/models/models_graph.py:17,23; registry train_gnn_embeddings.py:96-98
class SAGEConv(MessagePassing):
    def __init__(self, in_channels, out_channels, normalize=False, root_weight=True, bias=True,
                 aggr='mean'):
        super().__init__()
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.aggr = aggr
        self.normalize = normalize
        self.root_weight = root_weight
        self.lin_l = Linear(in_channels[0], out_channels, bias=bias)
        if root_weight:
            self.lin_r = Linear(in_channels[1], out_channels, bias=False)

    def reset_parameters(self):
        self.lin_l.reset_parameters()
        if self.root_weight:
            self.lin_r.reset_parameters()

    def forward(self, x, edge_index):
        if torch.is_tensor(x):
            x = (x, x)
        out = propagate(x[0], edge_index, x[1].shape[0], self.aggr)
        out = self.lin_l(out)
        if self.root_weight:
            out = out + self.lin_r(x[1])
        if self.normalize:
            out = F.normalize(out, p=2.0, dim=-1)
        return out


# a-5  GraphConv -- same slot; lin_rel (bias) / lin_root (no bias), neighbour aggregation 'add'
class GraphConv(MessagePassing):
    def __init__(self, in_channels, out_channels, aggr='add', bias=True):
        super().__init__()
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.aggr = aggr
        self.lin_rel = Linear(in_channels[0], out_channels, bias=bias)
        self.lin_root = Linear(in_channels[1], out_channels, bias=False)

    # PyG 1.7 names, also used by BASELINE.json's north_star
    @property
    def lin_l(self):
        return self.lin_rel

    @property
    def lin_r(self):
        return self.lin_root

    def reset_parameters(self):
        self.lin_rel.reset_parameters()
        self.lin_root.reset_parameters()

    def forward(self, x, edge_index):
        if torch.is_tensor(x):
            x = (x, x)
        out = propagate(x[0], edge_index, x[1].shape[0], self.aggr)
        out = self.lin_rel(out)
        return out + self.lin_root(x[1])


# 8f rank 3  GATConv -- the DEFAULT --operator of src/train_gnn_embeddings.py:15, registry :99.
# [PyG-recall] PyG 2.0.2 GATConv(in=(-1,-1), out, heads=1, concat=True, negative_slope=0.2,
# dropout=0.0, add_self_loops=True, bias=True):
#   x_l = lin_l(x_src), x_r = lin_r(x_dst)        (Linear, no bias; separate for tuple in_channels)
#   a_l = (x_l * att_l).sum(-1), a_r = (x_r * att_r).sum(-1)
#   edge_index: self loops removed, then (i, i) for i < min(N_src, N_dst) appended -- also for
#   bipartite edge types under to_hetero (the index pairs a source and a destination node of
#   different types; kept as the reference library does it)
#   e_ij = leaky_relu(a_l[j] + a_r[i]);  alpha = softmax over the incoming edges of i
#   (torch_geometric.utils.softmax: exp(e - max) / (sum + 1e-16));  out_i = sum_j alpha_ij x_l[j] + bias
def gat_edges(edge_index: torch.Tensor, n_src: int, n_dst: int) -> torch.Tensor:
    keep = edge_index[0] != edge_index[1]
    n = min(int(n_src), int(n_dst))
    loops = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index[:, keep], torch.stack([loops, loops])], dim=1)


class GATConv(MessagePassing):
    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2,
                 dropout=0.0, add_self_loops=True, bias=True):
        super().__init__()
        assert heads == 1 and concat and dropout == 0.0, 'oracle restates the reference defaults'
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.negative_slope, self.add_self_loops = negative_slope, add_self_loops
        self.out_channels = out_channels
        self.lin_l = Linear(in_channels[0], out_channels, bias=False, weight_initializer='glorot')
        self.lin_r = Linear(in_channels[1], out_channels, bias=False, weight_initializer='glorot')
        self.att_l = nn.Parameter(torch.empty(1, 1, out_channels))
        self.att_r = nn.Parameter(torch.empty(1, 1, out_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        self.lin_l.reset_parameters()
        self.lin_r.reset_parameters()
        nn.init.xavier_uniform_(self.att_l)          # PyG glorot
        nn.init.xavier_uniform_(self.att_r)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x, edge_index):
        if torch.is_tensor(x):
            x = (x, x)
        x_src, x_dst = x
        n_src, n_dst = x_src.shape[0], x_dst.shape[0]
        x_l = self.lin_l(x_src)
        x_r = self.lin_r(x_dst)
        a_l = (x_l * self.att_l.view(1, -1)).sum(-1)
        a_r = (x_r * self.att_r.view(1, -1)).sum(-1)
        ei = gat_edges(edge_index, n_src, n_dst) if self.add_self_loops else edge_index
        src, dst = ei[0], ei[1]
        e = F.leaky_relu(a_l[src] + a_r[dst], self.negative_slope)
        m = torch.full((n_dst,), float('-inf'), dtype=e.dtype).scatter_reduce(
            0, dst, e, reduce='amax', include_self=True)
        ex = torch.exp(e - m[dst])
        den = torch.zeros(n_dst, dtype=e.dtype).scatter_add_(0, dst, ex)
        alpha = ex / (den[dst] + 1e-16)
        msg = x_l.index_select(0, src) * alpha.view(-1, 1)
        out = torch.zeros(n_dst, self.out_channels, dtype=x_l.dtype).scatter_add_(
            0, dst.view(-1, 1).expand_as(msg), msg)
        if self.bias is not None:
            out = out + self.bias
        return out


# --------------------------------------------------------------------------------------------
# a-3  to_hetero group aggregation: pairwise torch.add through a FIFO queue, metadata edge order
# --------------------------------------------------------------------------------------------
def group_sum(outs_in_metadata_order):
    q = deque(outs_in_metadata_order)
    while len(q) >= 2:
        a, b = q.popleft(), q.popleft()
        q.append(torch.add(a, b))
    return q[0]


def hetero_conv(convs: nn.ModuleDict, x_dict, edge_index_dict, edge_types):
    per_dst = defaultdict(list)
    for et in edge_types:
        src, _, dst = et
        per_dst[dst].append(convs[key2str(et)]((x_dict[src], x_dict[dst]), edge_index_dict[et]))
    return OrderedDict((dst, group_sum(v)) for dst, v in per_dst.items())


# --------------------------------------------------------------------------------------------
# Hand-wired restatement of to_hetero(HeteroGNN) -- src/models/models_graph.py:5-49 as traced
# (SURVEY.md section 3.2): the pre-activation x flows to the next conv and is the embedding;
# relu -> dropout(training=True baked in by fx) only feeds conv_out.
# --------------------------------------------------------------------------------------------
class HeteroGNNOracle(nn.Module):
    def __init__(self, operator, activation, hidden_channels, out_channels, metadata, num_layers,
                 dropout, bn, skip):
        super().__init__()
        self.node_types, self.edge_types = metadata
        self.dropout, self.bn, self.skip = dropout, bn, skip
        self.traced_training = True          # modules are in training mode when to_hetero traces
        self.dropout_masks: Optional[Dict[str, torch.Tensor]] = None   # test hook: injected masks
        self.convs, self.lins, self.bns = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for _ in range(num_layers):
            self.convs.append(nn.ModuleDict(
                {key2str(et): operator((-1, -1), hidden_channels) for et in self.edge_types}))
            if skip:
                self.lins.append(nn.ModuleDict(
                    {t: Linear(-1, hidden_channels) for t in self.node_types}))
            else:
                self.lins.append(Linear(-1, hidden_channels))     # never called, stays lazy
            if bn:
                self.bns.append(nn.ModuleDict(
                    {t: nn.BatchNorm1d(hidden_channels) for t in self.node_types}))
            else:
                self.bns.append(nn.BatchNorm1d(hidden_channels))  # never called: not duplicated
        self.activation = nn.ModuleDict({t: copy.deepcopy(activation) for t in self.node_types})
        self.conv_out = nn.ModuleDict(
            {key2str(et): operator((-1, -1), out_channels) for et in self.edge_types})

    def forward(self, x, edge_index):
        x_emb = None
        for i, convs in enumerate(self.convs):
            h = hetero_conv(convs, x, edge_index, self.edge_types)
            if self.skip:
                h = OrderedDict((t, h[t] + self.lins[i][t](x[t])) for t in h)
            x = h
            if self.bn:
                x = OrderedDict((t, self.bns[i][t](v)) for t, v in x.items())
            x_emb = OrderedDict((t, self.activation[t](v)) for t, v in x.items())
            if self.traced_training:
                if self.dropout_masks is not None and i == len(self.convs) - 1:
                    x_emb = OrderedDict((t, v * self.dropout_masks[t]) for t, v in x_emb.items())
                elif self.dropout_masks is None:
                    x_emb = OrderedDict((t, F.dropout(v, self.dropout, training=True))
                                        for t, v in x_emb.items())
        x_out = hetero_conv(self.conv_out, x_emb, edge_index, self.edge_types)
        return x, OrderedDict((t, F.log_softmax(v, dim=1)) for t, v in x_out.items())


class HeteroSGNNOracle(nn.Module):
    """Same constructor as the reference's HeteroSGNN (src/models/models_graph.py:41-49)."""

    def __init__(self, operator, activation, aggr, hidden_channels, out_channels, metadata,
                 n_layers, dropout, bn, skip):
        super().__init__()
        assert aggr == 'sum'
        self.gnn = HeteroGNNOracle(operator, activation, hidden_channels, out_channels, metadata,
                                   n_layers, dropout, bn, skip)

    def forward(self, x, edge_index):
        emb, out_soft = self.gnn(x, edge_index)
        return emb, [out_soft]


# loss / accuracy -- src/train_gnn_embeddings.py:20-37
def nll_loss_artwork(out_soft, y_float):
    return F.nll_loss(out_soft['artwork'], y_float.type(torch.LongTensor))


def accuracy(predicted, labels):
    return predicted.argmax(dim=1).eq(labels).sum() / predicted.shape[0]


# --------------------------------------------------------------------------------------------
# Generic fx-based to_hetero (PyG 2.0.2 ``to_hetero_transformer`` algorithm), used by
# tests/golden/make_golden.py to run the reference's UNMODIFIED models_graph.py.
# --------------------------------------------------------------------------------------------
class _HeteroFx(nn.Module):
    def __init__(self, module: nn.Module, metadata, aggr: str = 'sum'):
        super().__init__()
        import torch.fx as fx
        assert aggr == 'sum'
        self.node_types, self.edge_types = metadata

        class _Tracer(fx.Tracer):
            def is_leaf_module(self, m, qualname):
                return isinstance(m, MessagePassing) or isinstance(m, Linear) or \
                    super().is_leaf_module(m, qualname)

        graph = _Tracer().trace(module)
        self._graph = graph
        called = {n.target for n in graph.nodes if n.op == 'call_module'}
        for target in called:
            sub = module.get_submodule(target)
            keys = ([key2str(et) for et in self.edge_types] if isinstance(sub, MessagePassing)
                    else list(self.node_types))
            md = nn.ModuleDict()
            for k in keys:
                md[k] = copy.deepcopy(sub)
                if hasattr(md[k], 'reset_parameters'):
                    md[k].reset_parameters()
            parent_name, _, leaf = target.rpartition('.')
            parent = module.get_submodule(parent_name) if parent_name else module
            if isinstance(parent, (nn.ModuleList, nn.Sequential)):
                parent[int(leaf)] = md
            else:
                setattr(parent, leaf, md)
        # like PyG's GraphModule, expose the transformed children at top level (state-dict keys
        # 'convs.0.<src>__<rel>__<dst>.lin_l.weight', 'bns.0.<type>.weight', ...)
        for name, child in module.named_children():
            self.add_module(name, child)
        for name, p in module.named_parameters(recurse=False):
            self.register_parameter(name, p)

    @property
    def root(self):
        return self

    def forward(self, x, edge_index):
        import torch.fx as fx
        env = {}
        placeholders = [n for n in self._graph.nodes if n.op == 'placeholder']
        env[placeholders[0]] = x
        env[placeholders[1]] = edge_index

        def load(a, key=None):
            def f(n):
                v = env[n]
                return v[key] if (key is not None and isinstance(v, dict)) else v
            return fx.node.map_arg(a, f)

        def dict_keys(args, kwargs):
            for a in list(args) + list(kwargs.values()):
                found = []
                fx.node.map_arg(a, lambda n: found.append(env[n]))
                for v in found:
                    if isinstance(v, dict):
                        return list(v.keys())
            return None

        for node in self._graph.nodes:
            if node.op == 'placeholder':
                continue
            if node.op == 'get_attr':
                env[node] = self.root.get_parameter(node.target)
            elif node.op == 'call_module':
                sub = self.root.get_submodule(node.target)
                first = next(iter(sub.values()))
                if isinstance(first, MessagePassing):
                    xd, eid = env[node.args[0]], env[node.args[1]]
                    env[node] = hetero_conv(sub, xd, eid, self.edge_types)
                else:
                    keys = dict_keys(node.args, node.kwargs)
                    env[node] = OrderedDict(
                        (k, sub[key2str(k)](*load(node.args, k), **load(node.kwargs, k)))
                        for k in keys)
            elif node.op in ('call_function', 'call_method'):
                keys = dict_keys(node.args, node.kwargs)

                def run(k):
                    a, kw = load(node.args, k), load(node.kwargs, k)
                    if node.op == 'call_function':
                        return node.target(*a, **kw)
                    return getattr(a[0], node.target)(*a[1:], **kw)
                env[node] = run(None) if keys is None else OrderedDict((k, run(k)) for k in keys)
            elif node.op == 'output':
                return load(node.args[0])


def to_hetero(module, metadata, aggr='sum'):
    return _HeteroFx(module, metadata, aggr)

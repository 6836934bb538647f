"""``to_hetero(module, metadata, aggr)``: turns a homogeneous GNN module into a heterogeneous one
(/root/reference/src/models/models_graph.py:45; PyG 2.0.2 ``to_hetero_transformer`` semantics,
SURVEY.md a-3) and executes it with fused agx launches:

  * every MessagePassing call site becomes ONE fused hetero layer over all edge types
    (``functional._HeteroConvFn``) instead of one conv call + ``torch.add`` chain per relation;
  * ``BatchNorm1d -> ReLU -> F.dropout`` per node type becomes one batched launch sequence over all
    node types; values nobody consumes (the reference's dead first-layer relu/dropout,
    models_graph.py:34-37) are not computed;
  * ``F.log_softmax(dim=1)`` per node type runs on the agx softmax kernel.

Leaf modules are duplicated exactly like PyG does (``ModuleDict`` keyed ``src__rel__dst`` for
MessagePassing modules, by node type for the others, ``reset_parameters()`` on each copy), so
state-dict keys match the reference: ``convs.0.artist__field_rel__field.lin_l.weight``,
``bns.0.artwork.running_mean``, ...
"""
from __future__ import annotations

import copy
import warnings
from collections import OrderedDict
from typing import Dict, List, Optional

import torch
import torch.fx as fx
import torch.nn as nn
import torch.nn.functional as F

from . import functional as AF
from . import ops
from . import _lib as L
from .functional import BNSpec, ConvSpec, RelSpec
from .data import Identity, identity_tensor, is_declared_identity
from .graph import get_plan, key2str
from .nn import Linear, MessagePassing


class _Tracer(fx.Tracer):
    def is_leaf_module(self, m, qualname):
        return isinstance(m, (MessagePassing, Linear)) or super().is_leaf_module(m, qualname)


def _is_dropout_fn(node: fx.Node) -> bool:
    return node.op == 'call_function' and node.target in (F.dropout, torch.dropout)


def _is_log_softmax(node: fx.Node) -> bool:
    return (node.op == 'call_function' and node.target in (F.log_softmax, torch.log_softmax)) or \
        (node.op == 'call_method' and node.target == 'log_softmax')


_IDENTITY_CACHE: Dict[tuple, bool] = {}


def _is_identity_input(x: torch.Tensor) -> bool:
    """One-hot node features (``torch.eye(N)``, src/data/artgraph.py:93-95) are detected once per
    tensor so that X @ W^T can be replaced by a transposed copy of W (exact)."""
    if x.dim() != 2 or x.shape[0] != x.shape[1] or x.shape[0] < 2:
        return False
    if is_declared_identity(x):          # data.Identity(n): declared by the caller, nothing to check
        return True
    key = (x.data_ptr(), x.shape[0], x._version)
    hit = _IDENTITY_CACHE.get(key)
    if hit is None:
        hit = bool(ops.is_identity(x).item())
        if len(_IDENTITY_CACHE) > 256:
            _IDENTITY_CACHE.clear()
        _IDENTITY_CACHE[key] = hit
    return hit


class HostDict(dict):
    """Result dict of a ``host_io`` module: values are computed on the device and copied to the
    host (through autograd, so ``backward`` carries the gradient back) the first time they are
    READ -- the reference reads ``out['artwork']`` only (train_gnn_embeddings.py:35), so the
    log-probabilities of the eight dead node types never cross PCIe."""

    def __init__(self, dev_dict):
        super().__init__((k, None) for k in dev_dict)
        self._dev = dev_dict
        self._host = {}

    def device_value(self, key):
        return self._dev[key]

    def __getitem__(self, key):
        if key not in self._host:
            self._host[key] = self._dev[key].cpu()
        return self._host[key]

    def get(self, key, default=None):
        return self[key] if key in self._dev else default

    def values(self):
        return [self[k] for k in self._dev]

    def items(self):
        return [(k, self[k]) for k in self._dev]


def _to_host(res):
    if isinstance(res, dict):
        return HostDict(res)
    if isinstance(res, (tuple, list)):
        return type(res)(_to_host(v) for v in res)
    return res.cpu() if torch.is_tensor(res) else res


class HeteroModule(nn.Module):
    def __init__(self, module: nn.Module, metadata, aggr: str = 'sum', detect_identity: bool = True,
                 host_io: bool = False):
        super().__init__()
        # host_io: accept the HOST tensors the reference's CPU-only script passes
        # (train_gnn_embeddings.py:42 -- it never calls .to(device)): inputs are staged on the
        # device once per tensor version, the module moves itself there on its first call, and
        # the result dicts hand values back to the host lazily (HostDict)
        self.host_io = bool(host_io)
        self._host_cache: Dict[tuple, torch.Tensor] = {}
        if aggr != 'sum':
            raise NotImplementedError("to_hetero: only aggr='sum' (the reference's setting, "
                                      "src/train_gnn_embeddings.py:130) is implemented")
        self.node_types, self.edge_types = list(metadata[0]), [tuple(e) for e in metadata[1]]
        self.detect_identity = detect_identity
        graph = _Tracer().trace(module)
        self._graph = graph
        called = []
        for n in graph.nodes:
            if n.op == 'call_module' and n.target not in called:
                called.append(n.target)
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
        for name, child in module.named_children():
            self.add_module(name, child)
        for name, p in module.named_parameters(recurse=False):
            self.register_parameter(name, p)
        # analysis results are keyed by node NAME so that copy.deepcopy(model) (the reference's
        # save_embeddings, src/train_gnn_embeddings.py:84-85) stays consistent
        self._placeholders = [n.name for n in graph.nodes if n.op == 'placeholder']
        self._live = self._liveness()
        self._fusion = self._plan_fusion()
        self._conv_specs: Dict[tuple, ConvSpec] = {}
        self._warned = set()
        self.dropout_masks: Optional[Dict[str, torch.Tensor]] = None   # test hook (injected masks)
        self._seed = None
        self._seed_common = None
        self._dist = None            # dist.DistContext: this module runs one rank of a multi-GPU job
        self._nbt_flat: dict = {}

    def __getstate__(self):
        # copy.deepcopy(model) (the reference's save_embeddings, train_gnn_embeddings.py:84-85) and
        # pickling: the cached layer plans hold process groups (not copyable) and every CSR / CSC
        # tensor of the graph; they are rebuilt on the next forward
        state = self.__dict__.copy()
        state['_conv_specs'] = {}
        state['_nbt_flat'] = {}
        state['_host_cache'] = {}
        return state

    def _stage(self, t: torch.Tensor, dev: torch.device) -> torch.Tensor:
        """Device copy of a host input tensor, made once per (storage, version)."""
        if t.device == dev:
            return t
        key = (t.data_ptr(), tuple(t.shape), t.dtype)
        hit = self._host_cache.get(key)
        if hit is None or hit[0] != t._version or hit[2] is not t:
            if len(self._host_cache) > 256:
                self._host_cache.clear()
            hit = (t._version, t.to(dev), t)       # (keeps t alive: data_ptr stays unique)
            self._host_cache[key] = hit
        return hit[1]

    def set_distributed(self, ctx):
        """Run as one rank of a destination-partitioned multi-GPU job (dist.DistContext):
        BatchNorm statistics over the rows of all ranks, boundary-row exchange before every conv,
        a different dropout stream per rank.  ``None`` restores single-GPU behaviour."""
        self._dist = ctx
        self._seed = None
        self._seed_common = None
        return self

    # ---- static analysis ----------------------------------------------------------------------
    def _liveness(self):
        live = set()
        out = [n for n in self._graph.nodes if n.op == 'output'][0]
        stack = [out]
        while stack:
            n = stack.pop()
            if n in live:
                continue
            live.add(n)
            stack.extend(n.all_input_nodes)
        return {n.name for n in live}

    def _plan_fusion(self):
        """BatchNorm1d call -> (ReLU module call -> (F.dropout call)) chains whose intermediate
        values have no other live consumer are computed by the BN kernel's epilogue."""
        fusion = {}
        for n in self._graph.nodes:
            if n.op != 'call_module' or n.name not in self._live:
                continue
            sub = next(iter(self.get_submodule(n.target).values()))
            if not isinstance(sub, nn.BatchNorm1d):
                continue
            relu = [u for u in n.users if u.name in self._live and u.op == 'call_module' and
                    isinstance(next(iter(self.get_submodule(u.target).values())), nn.ReLU)]
            if len(relu) != 1:
                continue
            relu = relu[0]
            live_users = [u for u in relu.users if u.name in self._live]
            drop = None
            if len(live_users) == 1 and _is_dropout_fn(live_users[0]):
                d = live_users[0]
                training = d.kwargs.get('training', d.args[2] if len(d.args) > 2 else True)
                if training is True and d.args[0] is relu:
                    p = d.kwargs.get('p', d.args[1] if len(d.args) > 1 else 0.5)
                    drop = (d.name, float(p))
            # y itself is written only if something else reads it (the model returning the last
            # hidden BatchNorm output as the embedding)
            y_live = any(u.name in self._live and u is not relu for u in n.users)
            fusion[n.name] = (relu.name, drop, y_live)
        return fusion

    # ---- execution ----------------------------------------------------------------------------
    def _seed_state(self, device, common: bool = False):
        """Philox (key, counter) on the device: one stream per rank, and -- ``common`` -- one that is
        the same on every rank (same torch seed everywhere) for the replicated node types."""
        if common:
            if getattr(self, '_seed_common', None) is None or self._seed_common.device != device:
                s = (torch.initial_seed() ^ 0x5DEECE66D) & 0x7fffffffffffffff
                st = torch.tensor([s, 0], dtype=torch.int64, device=device)
                if self._dist is not None:       # the ranks' torch seeds may differ: rank 0's
                    from .dist import broadcast_
                    broadcast_(st, self._dist.group)
                self._seed_common = st
            return self._seed_common
        if self._seed is None or self._seed.device != device:
            s = torch.initial_seed()
            if self._dist is not None:
                s += 0x9E3779B97F4A7C15 * (self._dist.rank + 1)
            s &= 0x7fffffffffffffff
            self._seed = torch.tensor([s, 0], dtype=torch.int64, device=device)
        return self._seed

    def _mask(self, shape, p, device, t):
        if self.dropout_masks is not None:
            return self.dropout_masks[t]
        st = self._seed_state(device)
        m = ops.dropout_mask(shape, p, st)
        n4 = (m.numel() + 3) // 4
        st[1] += n4                  # advance the Philox counter (device side: graph-capturable)
        return m

    def _masks(self, shapes, p, device, types):
        """Dropout masks of all node types from ONE Philox launch (one flat buffer, per-type
        views; every row is a multiple of 4 elements apart, so 128-bit accesses stay aligned)."""
        if self.dropout_masks is not None:
            return [self.dropout_masks[t] for t in types]
        rep = self._dist.replicated if self._dist is not None else frozenset()
        out = [None] * len(types)
        # replicated node types must see the SAME mask on every rank (their rows are computed by
        # all ranks and have to stay identical): a second stream whose seed does not depend on the
        # rank; the partitioned types draw from the per-rank stream
        for common in (False, True):
            sub = [i for i, t in enumerate(types) if (t in rep) == common]
            if not sub:
                continue
            st = self._seed_state(device, common=common)
            sizes = [((int(shapes[i][0]) * int(shapes[i][1]) + 3) // 4) * 4 for i in sub]
            flat = ops.dropout_mask((sum(sizes),), p, st)
            st[1] += sum(sizes) // 4     # advance the Philox counter (device side: graph-capturable)
            off = 0
            for i, sz in zip(sub, sizes):
                r, c = int(shapes[i][0]), int(shapes[i][1])
                out[i] = flat[off:off + r * c].view(r, c)
                off += sz
        return out

    def _virtual_masks(self, shapes, device, types):
        """What ``_masks`` would draw, without drawing it: per type the Philox state and the element
        offset of its mask inside the flat buffer agx_dropout_mask would have filled (the
        normalising kernel generates the same keep / drop decisions in place); plus the counter
        advances to apply once the kernels are queued."""
        rep = self._dist.replicated if self._dist is not None else frozenset()
        out = [None] * len(types)
        advance = []
        for common in (False, True):
            sub = [i for i, t in enumerate(types) if (t in rep) == common]
            if not sub:
                continue
            st = self._seed_state(device, common=common)
            off = 0
            for i in sub:
                out[i] = (st, off)
                off += ((int(shapes[i][0]) * int(shapes[i][1]) + 3) // 4) * 4
            advance.append((st, off // 4))
        return out, advance

    def _bump_batches_tracked(self, target, bns, types):
        """``num_batches_tracked += 1`` of every BatchNorm of a layer as one add: the per-module
        counters are re-pointed (once) at elements of one flat int64 tensor."""
        live = [bns[t] for t in types if bns[t].training and bns[t].track_running_stats and
                bns[t].num_batches_tracked is not None]
        if not live:
            return
        key = (target, tuple(id(b) for b in live))
        flat = self._nbt_flat.get(key)
        ok = flat is not None and all(
            b.num_batches_tracked.data_ptr() == flat[i].data_ptr() for i, b in enumerate(live))
        if not ok:
            flat = torch.stack([b.num_batches_tracked.detach().reshape(()) for b in live])
            for i, b in enumerate(live):
                b.num_batches_tracked = flat[i]
            self._nbt_flat = {key: flat} if len(self._nbt_flat) > 16 else {**self._nbt_flat, key: flat}
        flat += 1

    def _conv_gat(self, node, x_dict, ei_dict, is_input):
        """GATConv (the reference script's default --operator): every edge type of the layer in ONE
        fused call (functional._HeteroGATFn); relation outputs of a destination type are added in
        metadata order (PyG's pairwise torch.add queue differs only in summation order)."""
        convs = self.get_submodule(node.target)
        if self._dist is not None and (self._dist.halo is not None or self._dist.partial or
                                       self._dist.scatter):
            # (a graph block per rank needs no exchange inside the layer; a destination partition
            # that cuts edges would need GATConv's self loops in global node ids, and a softmax
            # over a row whose edges are spread over the ranks)
            raise NotImplementedError('GATConv on a partition that cuts edges is not implemented '
                                      '(SAGEConv / GraphConv are)')
        types = [t for t in self.node_types if t in x_dict]
        dev = x_dict[types[0]].device
        params: List[torch.Tensor] = []
        rels = []
        slope = None
        for et in self.edge_types:
            s, _, d = et
            conv = convs[key2str(et)]
            wl, wr, al, ar, b = conv.gat_params(x_dict[s].shape[1], x_dict[d].shape[1], dev)
            plan = AF.GATPlan.get(ei_dict[et], x_dict[s].shape[0], x_dict[d].shape[0],
                                  conv.add_self_loops)
            base = len(params)
            params += [wl, wr, al, ar]
            i_b = -1
            if b is not None:
                i_b = len(params)
                params.append(b)
            rels.append(AF.GATRelSpec(plan, s, d, base, base + 1, base + 2, base + 3, i_b))
            if slope is None:
                slope = float(conv.negative_slope)
            elif slope != float(conv.negative_slope):
                raise NotImplementedError('GATConv: one negative_slope per layer')
        spec = AF.GATSpec(node_types=types, rels=rels, out_channels=params[0].shape[0], slope=slope,
                          identity=({t: _is_identity_input(x_dict[t]) for t in types}
                                    if (is_input and self.detect_identity) else {}),
                          param_refs=params)
        outs = AF.hetero_gat(spec, [x_dict[t].contiguous() for t in types], params)
        return OrderedDict(zip(spec.dst_types, outs))

    def _conv(self, node, x_dict, ei_dict, plan, is_input):
        convs = self.get_submodule(node.target)
        if type(next(iter(convs.values()))).__name__ == 'GATConv':
            return self._conv_gat(node, x_dict, ei_dict, is_input)
        plan = plan()
        key = (node.target, id(plan))
        if self._dist is not None and self._dist.halo is not None:
            # boundary rows of the other ranks behind the owned rows (static inputs: once)
            x_dict = self._dist.halo.extend(x_dict, cache=is_input)
        types = [t for t in self.node_types if t in x_dict]
        dev = x_dict[types[0]].device
        params: List[torch.Tensor] = []
        rel_specs = []
        for et in self.edge_types:
            s, _, d = et
            conv = convs[key2str(et)]
            wl, bl, wr = conv.rel_params(x_dict[s].shape[1], x_dict[d].shape[1], dev)
            i_wl = len(params)
            params.append(wl)
            i_bl = i_wr = -1
            if bl is not None:
                i_bl = len(params)
                params.append(bl)
            if wr is not None:
                i_wr = len(params)
                params.append(wr)
            rel_specs.append((et, conv.aggr == 'mean', i_wl, i_bl, i_wr))
        spec = self._conv_specs.get(key)
        if spec is None:
            partial = (self._dist.partial or {}) if self._dist is not None else {}
            scatter = (self._dist.scatter or {}) if self._dist is not None else {}
            for et, cnt in list(partial.items()) + list(scatter.items()):
                # a partial relation's local rows hold SOME of a destination's edges: the divisor of
                # scatter-mean (and of its transpose) is the in-degree over the edges of all ranks
                plan[et].csr.cnt = cnt
            spec = ConvSpec(node_types=types,
                            rels=[RelSpec(plan[et], mean, a, b, c, partial=et in partial,
                                          own_rows=self._dist.n_owned[et[2]] if et in scatter else -1)
                                  for et, mean, a, b, c in rel_specs],
                            out_channels=params[0].shape[0],
                            group=self._dist.group if (partial or scatter) else None)
            if len(self._conv_specs) > 64:
                self._conv_specs.clear()
            self._conv_specs[key] = spec
        spec.identity = ({t: _is_identity_input(x_dict[t]) for t in types}
                         if (is_input and self.detect_identity) else {})
        spec.param_refs = params
        outs = AF.hetero_conv(spec, [x_dict[t].contiguous() for t in types], params)
        return OrderedDict(zip(spec.dst_types, outs))

    def _bn(self, node, x_dict):
        bns = self.get_submodule(node.target)
        relu, drop, y_live = self._fusion.get(node.name, (None, None, True))
        types = list(x_dict.keys())
        first = bns[types[0]]
        dmasks = None
        virt = None            # dropout without mask tensors: (seed state, offset) per type
        p_drop = 0.0
        if drop is not None:
            p = drop[1]
            if self.dropout_masks is not None:
                dmasks = self._masks([x_dict[t].shape for t in types], p,
                                     x_dict[types[0]].device, types)
            elif p > 0:
                p_drop = float(p)
                virt, advance = self._virtual_masks([x_dict[t].shape for t in types],
                                                    x_dict[types[0]].device, types)
        self._bump_batches_tracked(node.target, bns, types)

        def run(sub: List[int], grouped: bool):
            """BatchNorm of the node types ``sub`` (indices into ``types``) in one batched call;
            ``grouped``: statistics over the rows of all ranks."""
            tt = [types[i] for i in sub]
            spec = BNSpec(n=len(tt), F=first.num_features, training=first.training or
                          not first.track_running_stats, momentum=first.momentum, eps=first.eps,
                          running=[(bns[t].running_mean, bns[t].running_var) for t in tt],
                          with_act=relu is not None,
                          dmasks=None if dmasks is None else [dmasks[i] for i in sub],
                          drop=None if virt is None else [virt[i] for i in sub], drop_p=p_drop,
                          need_y=relu is None or y_live,
                          param_refs=([bns[t].weight for t in tt], [bns[t].bias for t in tt]))
            if grouped:
                spec.group = self._dist.group
                ctx, dev = self._dist, x_dict[tt[0]].device
                spec.counts = ctx.counts(tt, dev)
                spec.counts_of = lambda idx: ctx.counts([tt[i] for i in idx], dev)
            return AF.batch_norm_act(spec, [x_dict[t].contiguous() for t in tt],
                                     [bns[t].weight for t in tt], [bns[t].bias for t in tt])

        n = len(types)
        rep = self._dist.replicated if self._dist is not None else frozenset()
        if self._dist is None or not rep:
            res = run(list(range(n)), self._dist is not None)
        else:
            # replicated node types hold all their rows on every rank: their batch statistics are
            # local (and identical everywhere); the partitioned types' run over all ranks
            both = relu is not None and y_live
            res = [None] * (2 * n if both else n)
            for sub, grouped in (([i for i in range(n) if types[i] not in rep], True),
                                 ([i for i in range(n) if types[i] in rep], False)):
                if not sub:
                    continue
                out = run(sub, grouped)
                for j, i in enumerate(sub):
                    res[i] = out[j]
                    if both:
                        res[n + i] = out[len(sub) + j]
        n = len(types)
        if virt is not None:
            for st, k in advance:      # the kernels above read the counters: advance them now
                st[1] += k             # (device side: graph-capturable)
        if relu is not None and not y_live:      # fused: only relu(y) [* dropout] is written
            return None, OrderedDict(zip(types, res[:n])), relu, drop
        y = OrderedDict(zip(types, res[:n]))
        act = OrderedDict(zip(types, res[n:])) if relu is not None else None
        return y, act, relu, drop

    def _generic_warn(self, what):
        if what not in self._warned:
            self._warned.add(what)
            warnings.warn(f'mmac_b200.to_hetero: {what} runs through generic torch CUDA ops, not an '
                          f'agx kernel', stacklevel=3)

    def forward(self, x, edge_index):
        x_dict, ei_dict = x, edge_index
        if any(isinstance(v, Identity) for v in x_dict.values()):
            # declared one-hot features: a device-side eye tensor carries the shape, no upload
            ref = next((v for v in x_dict.values() if torch.is_tensor(v)), None)
            dev = ref.device if (ref is not None and ref.is_cuda) else L.compute_device()
            x_dict = OrderedDict((k, identity_tensor(v.n, dev) if isinstance(v, Identity) else v)
                                 for k, v in x_dict.items())
        to_host = False
        if self.host_io and any(not v.is_cuda for v in x_dict.values()):
            dev = L.compute_device()
            if any(p.device != dev for p in self.parameters()
                   if not isinstance(p, nn.parameter.UninitializedParameter)) or \
                    any(b.device != dev for b in self.buffers()):
                self.to(dev)        # in place: optimizers created earlier keep their references
            x_dict = OrderedDict((k, self._stage(v, dev)) for k, v in x_dict.items())
            ei_dict = OrderedDict((k, self._stage(v, dev)) for k, v in ei_dict.items())
            to_host = True
        num_nodes = {t: v.shape[0] for t, v in x_dict.items()}
        ei_dict = OrderedDict((tuple(k), v) for k, v in ei_dict.items())
        num_dst = None
        if self._dist is not None:
            num_nodes, num_dst = self._dist.plan_rows(num_nodes)
        plan_box: list = []

        def plan():
            # CSR / CSC of the plain edge lists (SAGEConv / GraphConv); GATConv layers sort their
            # own self-loop augmented lists (functional.GATPlan) and never ask for this one
            if not plan_box:
                plan_box.append(get_plan(OrderedDict((et, ei_dict[et]) for et in self.edge_types),
                                         num_nodes, num_dst=num_dst))
            return plan_box[0]
        env = {self._placeholders[0]: x_dict, self._placeholders[1]: ei_dict}
        provided = {}

        def load(a, key=None):
            def f(n):
                v = env[n.name]
                return v[key] if (key is not None and isinstance(v, dict)) else v
            return fx.node.map_arg(a, f)

        def dict_keys(args, kwargs):
            found = []
            fx.node.map_arg((args, kwargs), lambda n: found.append(env[n.name]))
            for v in found:
                if isinstance(v, dict):
                    return list(v.keys())
            return None

        for node in self._graph.nodes:
            if node.op == 'placeholder' or node.name not in self._live:
                continue
            if node.name in provided:
                env[node.name] = provided[node.name]
                continue
            if node.op == 'get_attr':
                env[node.name] = self.get_parameter(node.target)
            elif node.op == 'call_module':
                sub = self.get_submodule(node.target)
                first = next(iter(sub.values()))
                if isinstance(first, MessagePassing):
                    xin = env[node.args[0].name]
                    env[node.name] = self._conv(node, xin, ei_dict, plan,
                                                is_input=node.args[0].name == self._placeholders[0])
                elif isinstance(first, nn.BatchNorm1d):
                    y, act, relu, drop = self._bn(node, env[node.args[0].name])
                    env[node.name] = y
                    if relu is not None:
                        provided[drop[0] if drop is not None else relu] = act
                        if drop is not None:
                            provided[relu] = None      # consumed only by the fused dropout
                else:
                    if not isinstance(first, (Linear, nn.ReLU)):
                        self._generic_warn(type(first).__name__)
                    keys = dict_keys(node.args, node.kwargs)
                    env[node.name] = OrderedDict(
                        (k, sub[key2str(k)](*load(node.args, k), **load(node.kwargs, k)))
                        for k in keys)
            elif _is_log_softmax(node):
                src = env[node.args[0].name]
                keys = list(src.keys())
                env[node.name] = OrderedDict(zip(keys, AF.log_softmax_many([src[k] for k in keys])))
            elif _is_dropout_fn(node):
                src = env[node.args[0].name]
                p = node.kwargs.get('p', node.args[1] if len(node.args) > 1 else 0.5)
                training = node.kwargs.get('training', node.args[2] if len(node.args) > 2 else True)
                if not training or (p == 0 and self.dropout_masks is None):
                    env[node.name] = src
                else:
                    env[node.name] = OrderedDict(
                        (k, AF.dropout_apply(v, self._mask(v.shape, p, v.device, k)))
                        for k, v in src.items())
            elif node.op in ('call_function', 'call_method'):
                keys = dict_keys(node.args, node.kwargs)
                self._generic_warn(str(node.target))

                def run(k):
                    a, kw = load(node.args, k), load(node.kwargs, k)
                    if node.op == 'call_function':
                        return node.target(*a, **kw)
                    return getattr(a[0], node.target)(*a[1:], **kw)
                env[node.name] = run(None) if keys is None else OrderedDict((k, run(k)) for k in keys)
            elif node.op == 'output':
                res = load(node.args[0])
                return _to_host(res) if to_host else res
        raise RuntimeError('fx graph without output node')


def to_hetero(module: nn.Module, metadata, aggr: str = 'sum', **kwargs) -> HeteroModule:
    return HeteroModule(module, metadata, aggr, **kwargs)

"""Autograd functions over the agx kernels: the fused heterogeneous conv layer, batched
BatchNorm+ReLU+dropout, log_softmax / nll, the fusion-head linear and SmoothL1.

Arithmetic restated from (paths under /root/reference):
  SAGEConv / GraphConv inside to_hetero            src/models/models_graph.py:17,23,30,38,45
  BatchNorm1d / activation / dropout / log_softmax src/models/models_graph.py:19,32-39
  nll_loss                                         src/train_gnn_embeddings.py:29-30
  cat -> Dropout -> Linear, CE, SmoothL1           src/models/models_kg.py:237-243,
                                                   src/train_new_multimodal_multitask.py:79-81,
                                                   src/train_projector.py:33,52
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from . import ops
from ._lib import check, lib, ptr, stream_ptr
from .graph import HeteroPlan, Relation

# relative cost model for choosing transform-first vs aggregate-first per relation
_NO_INPLACE = __import__('os').environ.get('AGX_NO_INPLACE_GRADS') is not None   # A/B switch
_FLOP_RATE = 40e12      # sustained fp32 FFMA flop/s of the grouped GEMM (whole chip)
_CTA_FLOP_RATE = 0.25e12  # one CTA's share: latency floor of a problem with a single row tile
_BYTE_RATE = 5e12       # gather bandwidth


@dataclass
class RelSpec:
    rel: Relation
    mean: bool                    # SAGEConv: mean ; GraphConv: add
    i_wl: int                     # indices into the flat parameter list
    i_bl: int                     # -1: no bias
    i_wr: int                     # -1: no root weight
    transform_first: bool = False
    partial: bool = False         # multi-GPU: this rank holds SOME of each destination's edges
                                  # (partitioned source -> replicated destination type); the
                                  # neighbour sums of all ranks are all-reduced inside the layer
    own_rows: int = -1            # multi-GPU scatter relation (dist.GraphPartition(scattered=)): the
                                  # rank's edges give partial sums for the rel.n_dst = world * chunk
                                  # destination rows of ALL ranks; a reduce-scatter leaves the full
                                  # sums of the own_rows rows this rank owns (backward: all-gather)

    @property
    def rows(self) -> int:
        """Destination rows this rank computes the layer's output for."""
        return self.own_rows if self.own_rows >= 0 else self.rel.n_dst


@dataclass
class ConvSpec:
    """Static description of one hetero conv layer (all relations) for _HeteroConvFn."""
    node_types: List[str]         # types whose features are passed in, in tensor order
    rels: List[RelSpec]
    out_channels: int
    dst_types: List[str] = field(default_factory=list)   # output order
    identity: Dict[str, bool] = field(default_factory=dict)   # type -> x[type] is eye(N)
    planned: bool = False
    param_refs: Optional[list] = None    # the nn.Parameters behind the flat parameter list
    group: object = None                 # process group of the partial relations' all-reduce

    def plan_modes(self, feat_dims: Dict[str, int]):
        """Per relation: transform-first (Y = X_src W_l^T, gather at width O) or aggregate-first
        (gather at width F_src, product on the N_dst rows), whichever moves fewer bytes / flops.
        A GEMM with few output rows runs on few CTAs, so its cost is floored by the serial
        K-loop of one CTA (~_CTA_FLOP_RATE); one-hot sources make the transform a transposed
        copy of the weight."""
        O = self.out_channels
        for rs in self.rels:
            r = rs.rel
            fs = feat_dims[r.src]

            def gemm_cost(m):
                flops = 2.0 * m * fs * O
                return max(flops / _FLOP_RATE, 2.0 * min(m, 128) * fs * O / _CTA_FLOP_RATE)
            if self.identity.get(r.src, False):
                tf = r.n_src * O * 8.0 / _BYTE_RATE
            else:
                tf = gemm_cost(r.n_src)
            tf += r.n_edges * O * 4.0 / _BYTE_RATE
            af = r.n_edges * fs * 4.0 / _BYTE_RATE + gemm_cost(r.n_dst)
            rs.transform_first = tf < af or (tf <= af * 1.05 and r.n_src <= r.n_dst)
            if rs.partial or rs.own_rows >= 0:
                # aggregate first: the [N_dst, F_src] partial sums are what is all-reduced, before
                # any term that every rank computes in full (root weight, bias) is added
                rs.transform_first = False
        self.dst_types = []
        for rs in self.rels:
            if rs.rel.dst not in self.dst_types:
                self.dst_types.append(rs.rel.dst)
        self.planned = True


def _t(w: torch.Tensor) -> torch.Tensor:
    return w.t()


class _HeteroConvFn(torch.autograd.Function):
    """out[t] = sum_{r: dst(r)=t} ( lin_l_r(aggr_r(x[src(r)])) + lin_r_r(x[t]) )  for all t at once.

    Per relation the cheaper of transform-first (Y = X W_l^T on the small source table, then one
    fused multi-relation gather per destination row) and aggregate-first (gather at input width,
    then the product as one segment of the destination type's GEMM) is used; both equal the
    reference's aggregate-then-transform in exact arithmetic (fp32 drift ~1e-7).
    """

    @staticmethod
    def forward(ctx, spec: ConvSpec, *tensors):
        ctx.set_materialize_grads(False)      # dead outputs (conv_out of non-artwork types) stay None
        ctx.anchor = bool(getattr(spec, 'anchor', False))
        if ctx.anchor:                        # trailing dummy input, see _direct_call
            tensors = tensors[:-1]
        nt = len(spec.node_types)
        xs = {t: tensors[i] for i, t in enumerate(spec.node_types)}
        params = tensors[nt:]
        for t, x in xs.items():
            L.require_cuda(x, f'x[{t}]')
            if x.dtype != torch.float32 or not x.is_contiguous():
                raise TypeError(f'x[{t}] must be contiguous float32')
        if not spec.planned:
            spec.plan_modes({t: x.shape[1] for t, x in xs.items()})
        O = spec.out_channels
        dev = tensors[0].device
        by_dst: Dict[str, List[RelSpec]] = {t: [] for t in spec.dst_types}
        for rs in spec.rels:
            by_dst[rs.rel.dst].append(rs)
        xd = _dst_views(spec, xs)

        # s0: per destination type, the sum of root weights and of biases
        wroot: Dict[str, Optional[torch.Tensor]] = {}
        bsum: Dict[str, Optional[torch.Tensor]] = {}
        sums = []
        for t, lst in by_dst.items():
            roots = [params[rs.i_wr] for rs in lst if rs.i_wr >= 0]
            bias = [params[rs.i_bl] for rs in lst if rs.i_bl >= 0]
            for store, items in ((wroot, roots), (bsum, bias)):
                if not items:
                    store[t] = None
                elif len(items) == 1:
                    store[t] = items[0]
                else:
                    acc = torch.empty_like(items[0])
                    # at most 8 inputs per descriptor: chain
                    cur, rest = items[:8], items[8:]
                    sums.append((acc, cur))
                    while rest:
                        sums.append((acc, [acc] + rest[:7]))
                        rest = rest[7:]
                    store[t] = acc
        if sums:                    # independent sums in one launch, chained ones (> 8 inputs) after
            ops.sum_arrays([s for s in sums if s[1][0] is not s[0]])
            for s in sums:
                if s[1][0] is s[0]:
                    ops.sum_arrays([s])

        # s1: transform-first products on the source tables
        Y: Dict[int, torch.Tensor] = {}
        gb = ops.GemmBatch()
        tr: list = []                 # transposed weight copies, one batched launch
        for k, rs in enumerate(spec.rels):
            if rs.transform_first:
                r = rs.rel
                y = torch.empty(r.n_src, O, dtype=torch.float32, device=dev)
                Y[k] = y
                if spec.identity.get(r.src, False):
                    tr.append((y, params[rs.i_wl]))              # X = I  =>  X W^T = W^T exactly
                else:
                    gb.add(y, [(xs[r.src], _t(params[rs.i_wl]))])

        # s2: gathers
        outs: Dict[str, torch.Tensor] = {}
        has_tf: Dict[str, bool] = {}
        G: Dict[int, torch.Tensor] = {}
        rows_by_F: Dict[int, list] = {}
        chunks_by_F: Dict[int, list] = {}
        id_root: Dict[str, bool] = {}
        tf_long: Dict[str, list] = {}
        # all destination types' outputs are row blocks of ONE buffer (consumers that run the same
        # row-wise op on every type -- log_softmax -- then need a single launch)
        out_rows = [lst[0].rows for lst in by_dst.values()]
        out_all = torch.empty(sum(out_rows), O, dtype=torch.float32, device=dev)
        row0 = 0
        for t, lst in by_dst.items():
            outs[t] = out_all[row0:row0 + lst[0].rows]
            row0 += lst[0].rows
            # root product against one-hot features: x[t] W^T = W^T, written first, rest accumulates
            id_root[t] = bool(spec.identity.get(t, False)) and wroot[t] is not None and \
                xd[t] is xs[t]
            if id_root[t]:
                tr.append((outs[t], wroot[t]))
            tf_all = [(k, rs) for k, rs in enumerate(spec.rels)
                      if rs.rel.dst == t and rs.transform_first]
            has_tf[t] = bool(tf_all) or id_root[t]
            # transform-first relations with long / skewed rows: edge-balanced kernel into a
            # private [N_dst, O] buffer (N_dst is small there), summed into out[t] afterwards
            tf = [(k, rs) for k, rs in tf_all if not rs.rel.csr.long_rows]
            for k, rs in tf_all:
                if rs.rel.csr.long_rows:
                    tmp = torch.empty_like(outs[t])
                    tf_long.setdefault(t, []).append(tmp)
                    chunks_by_F.setdefault(O, []).append(
                        (tmp, ops.RelArg(rs.rel.csr, Y[k], mean_rows=rs.mean)))
            for base in range(0, len(tf), L.MAX_REL_PER_GROUP):
                part = tf[base:base + L.MAX_REL_PER_GROUP]
                # groups that accumulate onto the same output go to later launches ("waves")
                rows_by_F.setdefault((base // L.MAX_REL_PER_GROUP, O), []).append(
                    (outs[t], [ops.RelArg(rs.rel.csr, Y[k], mean_rows=rs.mean) for k, rs in part],
                     base > 0 or id_root[t]))
        part_k = [k for k, rs in enumerate(spec.rels) if rs.partial]
        part_flat = None
        if part_k:
            sizes = [spec.rels[k].rel.n_dst * xs[spec.rels[k].rel.src].shape[1] for k in part_k]
            part_flat = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
            off = 0
            for k, sz in zip(part_k, sizes):
                G[k] = part_flat[off:off + sz].view(spec.rels[k].rel.n_dst, -1)
                off += sz
        # aggregations of the partial relations run first: their all-reduce is started as soon as
        # they are done and overlaps everything of the layer that does not read the full sums
        part_rows: Dict[int, list] = {}
        part_chunks: Dict[int, list] = {}
        scat_full: Dict[int, torch.Tensor] = {}
        for k, rs in enumerate(spec.rels):
            if rs.transform_first:
                continue
            r = rs.rel
            fs = xs[r.src].shape[1]
            g = G[k] if rs.partial else torch.empty(r.n_dst, fs, dtype=torch.float32, device=dev)
            G[k] = g
            early = rs.partial or rs.own_rows >= 0
            if rs.own_rows >= 0:
                scat_full[k] = g
            arg = ops.RelArg(r.csr, xs[r.src], mean_rows=rs.mean)
            if r.csr.long_rows:
                (part_chunks if early else chunks_by_F).setdefault(fs, []).append((g, arg))
            elif early:
                part_rows.setdefault(fs, []).append((g, [arg], False))
            else:
                rows_by_F.setdefault((0, fs), []).append((g, [arg], False))
        works = []
        if part_flat is not None or scat_full:
            for F, grp in part_rows.items():
                ops.aggregate_rows(grp, F)
            for F, segs in part_chunks.items():
                ops.aggregate_chunks(segs, F)
            import torch.distributed as dist
            # scatter relations: every owner receives the full sums of its chunk of rows
            for k, full in scat_full.items():
                world = dist.get_world_size(spec.group)
                own = torch.empty(full.shape[0] // world, full.shape[1], dtype=torch.float32,
                                  device=dev)
                from .dist import reduce_scatter_rows
                works.append(reduce_scatter_rows(own, full, spec.group))
                G[k] = own[:spec.rels[k].own_rows]
            # partial neighbour sums of all ranks -> the full sums, on every rank (one all-reduce
            # per layer, on the communicator's stream; the mean relations were divided by the
            # GLOBAL in-degree above)
            if part_flat is not None:
                works.append(dist.all_reduce(part_flat, group=spec.group, async_op=True))
        # distinct outputs: the edge-balanced launches run on the side stream next to the row-parallel
        # ones; those that gather input features (not a transformed table Y) start before the
        # transform-first products
        y_ids = {id(y) for y in Y.values()}
        early = {F: [sg for sg in segs if id(sg[1].x) not in y_ids] for F, segs in chunks_by_F.items()}
        late = {F: [sg for sg in segs if id(sg[1].x) in y_ids] for F, segs in chunks_by_F.items()}
        n_early = sum(sg[1].csr.n_edges for segs in early.values() for sg in segs)
        n_late = sum(sg[1].csr.n_edges for segs in late.values() for sg in segs)
        if n_late == 0:
            early, late = chunks_by_F, {}           # nothing waits for the transform-first products
        elif min(n_early, n_late) < 0.25 * max(n_early, n_late, 1):
            # a second launch for a few small relations costs more than starting the others early
            early, late = {}, chunks_by_F
        fk = ops.fork(dev)
        with fk:
            for F, segs in early.items():
                ops.aggregate_chunks(segs, F)
        if tr:
            ops.transpose_many(tr)
        if gb.problems:
            gb.run()
        fk2 = ops.fork(dev)
        with fk2:
            for F, segs in late.items():
                ops.aggregate_chunks(segs, F)
        for (wave, F) in sorted(rows_by_F.keys()):
            ops.aggregate_rows(rows_by_F[(wave, F)], F)
        fk2.join()
        fk.join()
        if tf_long:
            items = []
            for t, temps in tf_long.items():
                already = id_root[t] or any(rs.transform_first and not rs.rel.csr.long_rows
                                            for rs in by_dst[t])
                ins = ([outs[t]] if already else []) + temps
                while len(ins) > 8:                       # chain (in place, elementwise)
                    ops.sum_arrays([(outs[t], ins[:8])])
                    ins = [outs[t]] + ins[8:]
                items.append((outs[t], ins))
            ops.sum_arrays(items)

        # s3: one multi-segment GEMM per destination type; the types that read all-reduced sums
        # wait for the collective, the others (artwork: most of the rows) are launched before
        part_types = {spec.rels[k].rel.dst for k in part_k} | \
            {spec.rels[k].rel.dst for k in scat_full}
        for dependent in (False, True):
            if dependent:
                for w in works:
                    if w is not None:
                        w.wait()
                scat_full.clear()
            gb = ops.GemmBatch()
            for t, lst in by_dst.items():
                if (t in part_types) != dependent:
                    continue
                segs = []
                for k, rs in enumerate(spec.rels):
                    if rs.rel.dst == t and not rs.transform_first:
                        segs.append((G[k], _t(params[rs.i_wl])))
                if wroot[t] is not None and not id_root[t]:
                    segs.append((xd[t], _t(wroot[t])))
                if not segs and bsum[t] is not None:
                    # bias only (no dense segment left): rank-1 product ones[N,1] @ bias[1,O]
                    ones = ops.fill_(torch.empty(outs[t].shape[0], 1, dtype=torch.float32, device=dev), 1.0)
                    gb.add(outs[t], [(ones, bsum[t].view(1, -1))], accumulate=has_tf[t])
                elif segs:
                    gb.add(outs[t], segs, bias=bsum[t], accumulate=has_tf[t])
                elif not has_tf[t]:
                    ops.fill_(outs[t], 0.0)
            if gb.problems:
                gb.run()

        ctx.spec = spec
        ctx.nt = nt
        ctx.g_keys = list(G.keys())
        ctx.save_for_backward(*tensors, *[G[k] for k in ctx.g_keys],
                              *[w for w in wroot.values() if w is not None])
        ctx.wroot_types = [t for t, w in wroot.items() if w is not None]
        return tuple(outs[t] for t in spec.dst_types)

    @staticmethod
    def backward(ctx, *douts):
        spec: ConvSpec = ctx.spec
        nt = ctx.nt
        saved = ctx.saved_tensors
        n_in = nt + _n_params(spec)
        tensors = saved[:n_in]
        xs = {t: tensors[i] for i, t in enumerate(spec.node_types)}
        params = tensors[nt:]
        G = {k: saved[n_in + i] for i, k in enumerate(ctx.g_keys)}
        wroot = {t: saved[n_in + len(G) + i] for i, t in enumerate(ctx.wroot_types)}
        xd = _dst_views(spec, xs)
        O = spec.out_channels
        dev = tensors[0].device
        dout: Dict[str, Optional[torch.Tensor]] = {}
        for t, g in zip(spec.dst_types, douts):
            dout[t] = None if g is None else g.contiguous()
        need_x = {t: ctx.needs_input_grad[1 + i] for i, t in enumerate(spec.node_types)}
        grads: List[Optional[torch.Tensor]] = [None] * len(tensors)
        pidx = lambda i: nt + i                                            # noqa: E731

        live = [(k, rs) for k, rs in enumerate(spec.rels) if dout[rs.rel.dst] is not None]

        # b0: d(full neighbour sum) of the partial relations: every rank holds a PART of that gradient
        # (its own consumers of the replicated rows); the transpose over a rank's edges needs the
        # sum over the ranks -> one flat buffer, one all-reduce, started first so that it overlaps
        # the rest of the layer's backward (it is only read by b5)
        dG: Dict[int, torch.Tensor] = {}
        part_k = [k for k, rs in live if rs.partial and need_x[rs.rel.src]]
        scat_k = [k for k, rs in live if rs.own_rows >= 0 and need_x[rs.rel.src]]
        dpart_flat = None
        works = []
        if part_k or scat_k:
            import torch.distributed as dist
            gb0 = ops.GemmBatch()
            # scatter relations: the owner computes d(full sum) of its rows; every rank needs all
            # rows for the transpose over its own edges -> all-gather
            d_own = {}
            for k in scat_k:
                rs = spec.rels[k]
                world = dist.get_world_size(spec.group)
                chunk, fs = rs.rel.n_dst // world, xs[rs.rel.src].shape[1]
                d_own[k] = torch.empty(chunk, fs, dtype=torch.float32, device=dev) \
                    if chunk == rs.own_rows else ops.zeros((chunk, fs), dev)
                gb0.add(d_own[k][:rs.own_rows], [(dout[rs.rel.dst], params[rs.i_wl])])
            sizes = [spec.rels[k].rel.n_dst * xs[spec.rels[k].rel.src].shape[1] for k in part_k]
            if part_k:
                dpart_flat = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
            off = 0
            for k, sz in zip(part_k, sizes):
                rs = spec.rels[k]
                dG[k] = dpart_flat[off:off + sz].view(rs.rel.n_dst, -1)
                off += sz
                gb0.add(dG[k], [(dout[rs.rel.dst], params[rs.i_wl])])
            gb0.run()
            for k in scat_k:
                rs = spec.rels[k]
                dG[k] = torch.empty(rs.rel.n_dst, d_own[k].shape[1], dtype=torch.float32, device=dev)
                works.append(dist.all_gather_into_tensor(dG[k], d_own[k], group=spec.group,
                                                         async_op=True))
            if part_k:
                works.append(dist.all_reduce(dpart_flat, group=spec.group, async_op=True))

        # b1: bias gradients = column sums of dout, shared by the relations of a type
        cs_items = []
        dbias: Dict[str, torch.Tensor] = {}
        for t in spec.dst_types:
            if dout[t] is not None and any(rs.i_bl >= 0 for _, rs in live if rs.rel.dst == t):
                dbias[t] = torch.empty(O, dtype=torch.float32, device=dev)
                cs_items.append((dout[t], dbias[t], False))
        # (launched below, on the side stream in front of the dY aggregation: nothing in this layer
        # reads the bias gradients)
        for k, rs in live:
            if rs.i_bl >= 0:
                grads[pidx(rs.i_bl)] = dbias[rs.rel.dst]

        # b4a: transform-first relations: dY = A^T dout  (transpose gather over the CSC)
        dY: Dict[int, torch.Tensor] = {}
        rows, chunks = [], []
        for k, rs in live:
            if not rs.transform_first:
                continue
            r = rs.rel
            dy = torch.empty(r.n_src, O, dtype=torch.float32, device=dev)
            dY[k] = dy
            arg = ops.RelArg(r.csc, dout[r.dst], mean_rows=False,
                             nbr_scale=r.csr.cnt if rs.mean else None)
            if r.csc.long_rows:
                chunks.append((dy, arg))
            else:
                rows.append((dy, [arg], False))
        # ... on the side stream, next to the products below that do not read dY (gathers are
        # latency / bandwidth bound, the products run on the tensor cores); the products that do
        # read dY (gb_late, trb_late) go with the last launch of the layer
        fk = ops.fork(dev)
        with fk:
            ops.colsum(cs_items)
            ops.aggregate_chunks(chunks, O)
            ops.aggregate_rows(rows, O)

        # b2/b3: weight gradients and the aggregate-first input gradients dG = dout W_l
        gb = ops.GemmBatch()          # products the rest of this backward reads (dG)
        gb_w = ops.GemmBatch()        # weight gradients from dout
        gb_late = ops.GemmBatch()     # weight gradients from dY
        trb: list = []
        trb_late: list = []
        dwroot: Dict[str, torch.Tensor] = {}
        # gradient buffers the optimizer already owns (FlatAdam's arena): a weight gradient with a
        # single producer is ADDED into its buffer by the producing kernel itself (the accumulate
        # forms of the GEMM epilogue / the transposed copy) instead of being written to a temporary
        # and added by _deliver_param_grads afterwards (one read + one write of every weight less;
        # the one-hot input layer holds most of the model's 5 M weights)
        refs = spec.param_refs
        in_place = refs is not None and not _NO_INPLACE and all(
            p.grad is not None and p.grad.is_contiguous() for p in refs)
        placed: Dict[int, bool] = {}

        def grad_buffer(i_param, shape):
            """(tensor to produce the gradient in, accumulate?)"""
            if in_place and tuple(refs[i_param].grad.shape) == tuple(shape):
                placed[i_param] = True
                return refs[i_param].grad, True
            return torch.empty(shape, dtype=torch.float32, device=dev), False

        for t in spec.dst_types:
            if dout[t] is None or t not in wroot:
                continue
            x = xd[t]
            owners = [rs.i_wr for _, rs in live if rs.rel.dst == t and rs.i_wr >= 0]
            if len(owners) == 1:                      # one relation's lin_r: straight into its buffer
                dw, acc = grad_buffer(owners[0], (O, x.shape[1]))
            else:
                dw, acc = torch.empty(O, x.shape[1], dtype=torch.float32, device=dev), False
            dwroot[t] = dw
            if spec.identity.get(t, False) and x is xs[t]:
                trb.append((dw, dout[t], acc))                   # dout^T I
            else:
                gb_w.add(dw, [(_t(dout[t]), x)], split_k=ops.split_k_for(x.shape[0]), accumulate=acc)
        for k, rs in live:
            r = rs.rel
            x = xs[r.src]
            dw, acc = grad_buffer(rs.i_wl, (O, x.shape[1]))
            grads[pidx(rs.i_wl)] = dw
            if rs.transform_first and spec.identity.get(r.src, False):
                trb_late.append((dw, dY[k], acc))
            elif rs.transform_first:
                gb_late.add(dw, [(_t(dY[k]), x)], split_k=ops.split_k_for(r.n_src), accumulate=acc)
            else:
                gb_w.add(dw, [(_t(dout[r.dst]), G[k])], split_k=ops.split_k_for(rs.rows),
                         accumulate=acc)
                if need_x[r.src] and k not in dG:
                    dg = torch.empty(r.n_dst, x.shape[1], dtype=torch.float32, device=dev)
                    dG[k] = dg
                    gb.add(dg, [(dout[r.dst], params[rs.i_wl])])
        # In direct-gradient mode nothing reads a weight gradient before the optimizer step: those
        # products leave the critical path (ops.defer: a third stream, joined when the whole
        # backward pass is over) and run under the BatchNorm backward / the next layer's backward.
        deferred = in_place and ctx.anchor
        hold = [t for t in dout.values() if t is not None] + list(xs.values()) + list(G.values()) + \
            list(dwroot.values())
        with (ops.defer(dev, hold) if deferred else ops.defer(torch.device('cpu'))) as dfr:
            if trb:
                ops.transpose_many(trb)
            if gb_w.problems:
                gb_w.run(dfr.keep)
        if gb.problems:
            gb.run()
        for w in works:
            w.wait()
        for k, rs in live:
            if rs.i_wr >= 0:
                grads[pidx(rs.i_wr)] = dwroot[rs.rel.dst]

        # b5: input gradients per source type
        groups = {}
        late_root: list = []
        long_chunks: Dict[int, list] = {}
        long_sums: list = []
        for t in spec.node_types:
            if not need_x[t]:
                continue
            af = [(k, rs) for k, rs in live if rs.rel.src == t and not rs.transform_first]
            tf = [(k, rs) for k, rs in live if rs.rel.src == t and rs.transform_first]
            root = dout.get(t) is not None and t in wroot
            if not (af or tf or root):
                continue
            dx = torch.empty_like(xs[t])
            grads[spec.node_types.index(t)] = dx
            fs = xs[t].shape[1]
            af_short = [(k, rs) for k, rs in af if not rs.rel.csc.long_rows]
            af_long = [(k, rs) for k, rs in af if rs.rel.csc.long_rows]
            for base in range(0, len(af_short), L.MAX_REL_PER_GROUP):
                part = af_short[base:base + L.MAX_REL_PER_GROUP]
                groups.setdefault((base // L.MAX_REL_PER_GROUP, fs), []).append(
                    (dx, [ops.RelArg(rs.rel.csc, dG[k], nbr_scale=rs.rel.csr.cnt if rs.mean else None)
                          for k, rs in part], base > 0))
            # skewed source rows (a source node feeding thousands of destinations): edge-balanced
            # kernel into a private buffer, summed into dx afterwards
            temps = []
            for k, rs in af_long:
                tmp = torch.empty_like(dx)
                temps.append(tmp)
                long_chunks.setdefault(fs, []).append(
                    (tmp, ops.RelArg(rs.rel.csc, dG[k], nbr_scale=rs.rel.csr.cnt if rs.mean else None)))
            if temps:
                long_sums.append((dx, ([dx] if af_short else []) + temps))
            segs = [(dY[k], params[rs.i_wl]) for k, rs in tf]
            if root and xd[t] is xs[t]:
                segs.append((dout[t], wroot[t]))
            elif root:
                # partitioned graph: the table carries other ranks' boundary rows behind the owned
                # rows; the root term only reaches the owned ones (second, accumulating launch)
                if not (af or tf):
                    ops.fill_(dx, 0.0)
                late_root.append((dx[:xd[t].shape[0]], [(dout[t], wroot[t])], True))
            groups.setdefault(('gemm',), []).append((dx, segs, bool(af)))
        fk5 = ops.fork(dev)
        with fk5:
            for F_, segs_ in long_chunks.items():
                ops.aggregate_chunks(segs_, F_)
        for key in sorted(k for k in groups.keys() if k != ('gemm',)):
            ops.aggregate_rows(groups[key], key[1])
        fk5.join()
        fk.join()                                   # dY is complete (b4a, same side stream)
        with (ops.defer(dev, list(dY.values())) if deferred
              else ops.defer(torch.device('cpu'))) as dfr:
            if trb_late:
                ops.transpose_many(trb_late)
            if gb_late.problems:
                gb_late.run(dfr.keep)
        last = []
        for dx_, ins in long_sums:
            while len(ins) > 8:
                ops.sum_arrays([(dx_, ins[:8])])
                ins = [dx_] + ins[8:]
            last.append((dx_, ins))
        if last:
            ops.sum_arrays(last)
        for dx, segs, acc in groups.get(('gemm',), []):
            if segs:
                gb.add(dx, segs, accumulate=acc)
        if gb.problems:
            gb.run()
        for dx, segs, acc in late_root:
            gb.add(dx, segs, accumulate=acc)
        if gb.problems:
            gb.run()
        for i_param in placed:                      # already added into the optimizer's buffers
            grads[pidx(i_param)] = None
        with (ops.defer(dev, [g for g in grads if g is not None]) if deferred
              else ops.defer(torch.device('cpu'))):
            _deliver_param_grads(spec.param_refs, grads, nt, must=ctx.anchor)
        return (None, *grads) + ((None,) if ctx.anchor else ())


def _dst_views(spec: ConvSpec, xs: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """x[t] as seen by the root (``lin_r``) term: the rows of the destination nodes.  Single GPU:
    the whole table; on a rank of a partitioned graph (dist.GraphPartition) the owned rows, which
    precede the gathered boundary rows."""
    out = {}
    for rs in spec.rels:
        t = rs.rel.dst
        if t in out or t not in xs:
            continue
        x = xs[t]
        if x.shape[0] < rs.rows:
            raise ValueError(f"x['{t}'] has {x.shape[0]} rows, the graph has {rs.rows}")
        out[t] = x if x.shape[0] == rs.rows else x[:rs.rows]
    return out


def _direct_call(fn, spec, x_list, params):
    """Call a fused layer function.  When every parameter of the layer already owns a gradient buffer
    (FlatAdam's flat arena) the parameters are passed DETACHED: the backward pass adds their
    gradients straight into those buffers (_deliver_param_grads) and autograd holds no edge to the
    parameters' AccumulateGrad nodes at all -- ~190 node visits per step less, and no dependence on
    the stream such a node was created on (a model that had run eagerly on the default stream
    could not be captured afterwards: "dependency created on uncaptured work in another stream").
    A fresh dummy leaf keeps the layer in the autograd graph when no feature input needs a gradient
    (the first conv layer)."""
    refs = spec.param_refs
    direct = torch.is_grad_enabled() and refs is not None and len(refs) == len(params) and all(
        isinstance(q, torch.nn.Parameter) and q.requires_grad and q.grad is not None and
        q.grad.is_contiguous() for q in refs)
    spec.anchor = direct
    if not direct:
        return fn.apply(spec, *x_list, *params)
    anchor = torch.empty((), dtype=torch.float32, device=x_list[0].device, requires_grad=True)
    return fn.apply(spec, *x_list, *[q.detach() for q in params], anchor)


def _deliver_param_grads(param_refs, grads, offset, must: bool = False):
    """When every parameter already owns a gradient buffer (FlatAdam's flat arena), add the fresh
    gradients into those buffers with ONE batched agx launch and hand autograd ``None`` -- instead
    of ~50 AccumulateGrad element-wise kernels per layer.  ``must``: the layer ran with detached
    parameters (_direct_call), so this is the only way the gradients reach them."""
    if param_refs is None:
        if must:
            raise RuntimeError('fused layer ran in direct-gradient mode without parameter references')
        return
    todo = [(i, p) for i, p in enumerate(param_refs) if grads[offset + i] is not None]
    if must and any(p.grad is None or not p.grad.is_contiguous() for _, p in todo):
        raise RuntimeError('a parameter gradient buffer disappeared between the forward and the '
                           'backward pass of a fused layer (use FlatAdam.zero_grad(), not '
                           'set_to_none)')
    if not todo or any(p.grad is None or not p.grad.is_contiguous() for _, p in todo):
        return
    items = []
    for i, p in todo:
        g = grads[offset + i]
        if not g.is_contiguous():
            return
        items.append((p.grad, [p.grad, g]))
    ops.sum_arrays(items)
    for i, _ in todo:
        grads[offset + i] = None


def _n_params(spec: ConvSpec) -> int:
    n = 0
    for rs in spec.rels:
        n = max(n, rs.i_wl + 1, rs.i_bl + 1, rs.i_wr + 1)
    return n


class _SageLayerFn(torch.autograd.Function):
    """One SAGEConv / GraphConv relation through the layer-level C entry points
    (agx_sage_layer_fwd / agx_sage_layer_bwd, include/agx.h): the path of a standalone operator call
    ``conv((x_src, x_dst), edge_index)``; inside ``to_hetero`` all relations of a layer run through
    the fused _HeteroConvFn instead."""

    @staticmethod
    def _layer(rel: Relation, mean: bool, x_src, x_dst, wl, bl, wr) -> "L.SageLayer":
        lay = L.SageLayer()
        r = lay.rel
        r.rowptr, r.col, r.cnt = ptr(rel.csr.rowptr), ptr(rel.csr.col), ptr(rel.csr.cnt)
        r.t_rowptr, r.t_col = ptr(rel.csc.rowptr), ptr(rel.csc.col)
        r.n_src, r.n_dst, r.n_edges = rel.n_src, rel.n_dst, rel.n_edges
        r.long_rows, r.t_long_rows = int(rel.csr.long_rows), int(rel.csc.long_rows)
        lay.mean = int(mean)
        lay.f_src, lay.out_channels = x_src.shape[1], wl.shape[0]
        lay.x_src, lay.ld_src = ptr(x_src), x_src.stride(0)
        if wr is not None:
            lay.f_dst = x_dst.shape[1]
            lay.x_dst, lay.ld_dst = ptr(x_dst), x_dst.stride(0)
        lay.w_l, lay.b_l, lay.w_r = ptr(wl), ptr(bl), ptr(wr)
        return lay

    @staticmethod
    def forward(ctx, rel: Relation, mean: bool, x_src, x_dst, wl, bl, wr):
        for t, name in ((x_src, 'x_src'), (x_dst, 'x_dst'), (wl, 'lin_l.weight')):
            L.require_cuda(t, name)
        x_src, x_dst = x_src.contiguous(), x_dst.contiguous()
        wl = wl.contiguous()
        wr = None if wr is None else wr.contiguous()
        if x_src.dtype != torch.float32 or x_dst.dtype != torch.float32:
            raise TypeError('node features must be float32')
        dev = x_src.device
        lay = _SageLayerFn._layer(rel, mean, x_src, x_dst, wl, bl, wr)
        nbytes = lib().agx_sage_layer_workspace_bytes(lay)
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
        out = torch.empty(rel.n_dst, wl.shape[0], dtype=torch.float32, device=dev)
        agg = torch.empty(rel.n_dst, x_src.shape[1], dtype=torch.float32, device=dev)
        check(lib().agx_sage_layer_fwd(lay, ptr(out), out.stride(0), 0, ptr(agg), ptr(ws), nbytes,
                                       stream_ptr()), 'agx_sage_layer_fwd')
        ctx.rel, ctx.mean = rel, mean
        ctx.has = (bl is not None, wr is not None)
        ctx.save_for_backward(x_src, x_dst, wl, agg, *([wr] if wr is not None else []))
        return out

    @staticmethod
    def backward(ctx, dout):
        x_src, x_dst, wl, agg, *rest = ctx.saved_tensors
        wr = rest[0] if rest else None
        has_b, has_r = ctx.has
        dout = dout.contiguous()
        dev = dout.device
        lay = _SageLayerFn._layer(ctx.rel, ctx.mean, x_src, x_dst, wl, None, wr)
        nbytes = lib().agx_sage_layer_workspace_bytes(lay)
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
        dwl = torch.empty_like(wl)
        dbl = torch.empty(wl.shape[0], dtype=torch.float32, device=dev) if has_b else None
        dwr = torch.empty_like(wr) if has_r else None
        dxs = torch.empty_like(x_src) if ctx.needs_input_grad[2] else None
        dxd = torch.empty_like(x_dst) if (ctx.needs_input_grad[3] and has_r) else None
        check(lib().agx_sage_layer_bwd(lay, ptr(agg), ptr(dout), dout.stride(0), ptr(dwl), ptr(dbl),
                                       ptr(dwr), ptr(dxs), x_src.shape[1], ptr(dxd),
                                       x_dst.shape[1] if has_r else 0, 0, ptr(ws), nbytes,
                                       stream_ptr()), 'agx_sage_layer_bwd')
        return None, None, dxs, dxd, dwl, dbl, dwr


def sage_layer(rel: Relation, mean: bool, x_src, x_dst, wl, bl=None, wr=None) -> torch.Tensor:
    return _SageLayerFn.apply(rel, mean, x_src, x_dst, wl, bl, wr)


def hetero_conv(spec: ConvSpec, x_list: Sequence[torch.Tensor], params: Sequence[torch.Tensor]):
    return _direct_call(_HeteroConvFn, spec, x_list, params)


# ------------------------------------------------------------------------------------------------
# BatchNorm1d (+ ReLU + dropout) batched over node types
# ------------------------------------------------------------------------------------------------
@dataclass
class BNSpec:
    n: int                                   # number of node types
    F: int
    training: bool
    momentum: float
    eps: float
    running: List[Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]]   # per type buffers
    with_act: bool                           # also produce relu(y) * dmask
    dmasks: Optional[List[Optional[torch.Tensor]]] = None
    # dropout without mask tensors: per type (seed state [2] int64, element offset into the virtual
    # flat mask agx_dropout_mask would have written) -- generated inside the normalising pass
    drop: Optional[List[Optional[Tuple[torch.Tensor, int]]]] = None
    drop_p: float = 0.0
    need_y: bool = True                      # False (with_act only): y itself is never written
    param_refs: Optional[tuple] = None       # ([weight Parameters], [bias Parameters])
    group: object = None                     # torch.distributed process group -> SyncBN
    counts: Optional[torch.Tensor] = None    # float64 [n] global row counts (with group)
    counts_of: Optional[object] = None       # callable(indices) -> counts of a subset (cached)


def _keep_scale(p: float) -> float:
    """1 / (1 - p) as the kernels compute it (float32 arithmetic)."""
    import numpy as np
    return float(np.float32(1.0) / (np.float32(1.0) - np.float32(p)))


class _BNActFn(torch.autograd.Function):
    """y_t = BatchNorm1d_t(x_t) for every node type t in one batched launch sequence; optionally
    also a_t = relu(y_t) * dmask_t (the input of conv_out, src/models/models_graph.py:34-38)."""

    @staticmethod
    def forward(ctx, spec: BNSpec, *tensors):
        ctx.set_materialize_grads(False)
        n, F = spec.n, spec.F
        xs, ws, bs = tensors[:n], tensors[n:2 * n], tensors[2 * n:3 * n]
        dev = xs[0].device
        arr = (L.BnDesc * n)()
        ys, acts, means, invstds = [], [], [], []
        rows = 0
        for i in range(n):
            x = xs[i]
            L.require_cuda(x, 'BatchNorm input')
            if x.dtype != torch.float32 or not x.is_contiguous() or x.shape[1] != F:
                raise TypeError('BatchNorm input must be contiguous float32 [N, F]')
            need_y = spec.need_y or not spec.with_act
            y = torch.empty_like(x) if need_y else None
            a = torch.empty_like(x) if spec.with_act else None
            m = torch.empty(F, dtype=torch.float32, device=dev)
            s = torch.empty(F, dtype=torch.float32, device=dev)
            dm = spec.dmasks[i] if (spec.with_act and spec.dmasks is not None) else None
            dr = spec.drop[i] if (spec.with_act and spec.drop is not None and dm is None) else None
            rm, rv = spec.running[i]
            arr[i] = L.BnDesc(ptr(x), ptr(y), ptr(a), ptr(dm), ptr(ws[i]), ptr(bs[i]), ptr(rm),
                              ptr(rv), ptr(m), ptr(s), x.shape[0],
                              float(spec.drop_p) if dr is not None else 0.0,
                              ptr(dr[0]) if dr is not None else None,
                              int(dr[1]) if dr is not None else 0)
            ys.append(y); acts.append(a); means.append(m); invstds.append(s)
            rows += x.shape[0]
        n_ws = lib().agx_bn_workspace_floats(rows, n, F)
        wsb = torch.empty(n_ws, dtype=torch.float32, device=dev)
        if spec.group is not None and spec.training:
            # SyncBN: statistics over the rows of ALL ranks (one float64 all-reduce of [n, 2F]:
            # column sums of x and of x*x)
            import torch.distributed as dist
            sums = torch.empty(n * 2 * F, dtype=torch.float64, device=dev)
            for phase in (1, 4):
                check(lib().agx_bn_forward_phase(arr, n, F, 1, spec.momentum, spec.eps, ptr(wsb),
                                                 n_ws, phase, ptr(sums), ptr(spec.counts),
                                                 stream_ptr()), 'agx_bn_forward_phase')
                if phase == 1:
                    from .dist import small_all_reduce_    # 18 KB: one peer-memory kernel, or NCCL
                    small_all_reduce_(sums, spec.group)
        else:
            check(lib().agx_bn_forward(arr, n, F, int(spec.training), spec.momentum, spec.eps,
                                       ptr(wsb), n_ws, stream_ptr()), 'agx_bn_forward')
        ctx.spec = spec
        # the relu gate of the backward pass: y, or y_act when y was not written (y_act > 0 <=> y > 0
        # and the element was kept; a dropped element gets no gradient either way)
        gates = [y if y is not None else a for y, a in zip(ys, acts)]
        ctx.save_for_backward(*xs, *ws, *gates, *means, *invstds)
        ctx.has_y = spec.need_y or not spec.with_act
        if not ctx.has_y:
            return tuple(acts)
        if spec.with_act:
            return (*ys, *acts)
        return tuple(ys)

    @staticmethod
    def backward(ctx, *grads):
        spec: BNSpec = ctx.spec
        n, F = spec.n, spec.F
        sv = ctx.saved_tensors
        xs, ws, ys, means, invstds = (sv[0:n], sv[n:2 * n], sv[2 * n:3 * n], sv[3 * n:4 * n],
                                      sv[4 * n:5 * n])
        if ctx.has_y:
            dys = grads[:n]
            dacts = grads[n:2 * n] if spec.with_act else [None] * n
        else:
            dys, dacts = [None] * n, grads[:n]
        dev = xs[0].device
        idx = [i for i in range(n) if dys[i] is not None or dacts[i] is not None]
        dxs: List[Optional[torch.Tensor]] = [None] * n
        dws: List[Optional[torch.Tensor]] = [None] * n
        dbs: List[Optional[torch.Tensor]] = [None] * n
        if idx:
            arr = (L.BnBwdDesc * len(idx))()
            rows = 0
            keep = []
            refs = spec.param_refs
            direct = refs is not None and all(
                refs[0][i].grad is not None and refs[1][i].grad is not None and
                refs[0][i].grad.is_contiguous() and refs[1][i].grad.is_contiguous() for i in idx)
            if not direct:
                zbuf = ops.zeros(2 * len(idx) * F, dev)          # the kernel accumulates (+=)
            for j, i in enumerate(idx):
                dy = None if dys[i] is None else dys[i].contiguous()
                da = None if dacts[i] is None else dacts[i].contiguous()
                keep += [dy, da]
                dxs[i] = torch.empty_like(xs[i]) if ctx.needs_input_grad[1 + i] else None
                if direct:      # straight into the optimizer's gradient arena, autograd gets None
                    dws[i], dbs[i] = refs[0][i].grad, refs[1][i].grad
                else:
                    dws[i] = zbuf[2 * j * F:(2 * j + 1) * F]
                    dbs[i] = zbuf[(2 * j + 1) * F:(2 * j + 2) * F]
                dm = spec.dmasks[i] if (spec.with_act and spec.dmasks is not None) else None
                philox = spec.with_act and dm is None and spec.drop is not None and \
                    spec.drop[i] is not None and spec.drop_p > 0
                arr[j] = L.BnBwdDesc(ptr(xs[i]), ptr(ys[i]), ptr(dy), ptr(da), ptr(dm), ptr(ws[i]),
                                     ptr(means[i]), ptr(invstds[i]), ptr(dxs[i]), ptr(dws[i]),
                                     ptr(dbs[i]), xs[i].shape[0],
                                     _keep_scale(spec.drop_p) if philox else 0.0)
                rows += xs[i].shape[0]
            n_ws = lib().agx_bn_workspace_floats(rows, len(idx), F)
            wsb = torch.empty(n_ws, dtype=torch.float32, device=dev)
            if spec.group is not None and spec.training:
                import torch.distributed as dist
                totals = torch.empty(len(idx) * 2 * F, dtype=torch.float64, device=dev)
                counts = spec.counts_of(tuple(idx)) if len(idx) != n else spec.counts
                check(lib().agx_bn_backward_phase(arr, len(idx), F, 1, ptr(wsb), n_ws, 1, ptr(totals),
                                                  ptr(counts), stream_ptr()), 'agx_bn_backward_phase')
                from .dist import small_all_reduce_
                small_all_reduce_(totals, spec.group)
                check(lib().agx_bn_backward_phase(arr, len(idx), F, 1, ptr(wsb), n_ws, 2, ptr(totals),
                                                  ptr(counts), stream_ptr()), 'agx_bn_backward_phase')
            else:
                check(lib().agx_bn_backward(arr, len(idx), F, int(spec.training), ptr(wsb), n_ws,
                                            stream_ptr()), 'agx_bn_backward')
            if direct:
                dws = [None] * n
                dbs = [None] * n
        return (None, *dxs, *dws, *dbs)


def batch_norm_act(spec: BNSpec, xs, weights, biases):
    refs = spec.param_refs
    direct = torch.is_grad_enabled() and refs is not None and all(
        q is not None and q.requires_grad and q.grad is not None and q.grad.is_contiguous()
        for q in (*refs[0], *refs[1]))
    if direct and any(x.requires_grad for x in xs):
        # gradients of weight / bias go straight into the parameters' buffers (see _direct_call);
        # the inputs keep the function in the autograd graph
        weights = [w.detach() for w in weights]
        biases = [b.detach() for b in biases]
    return _BNActFn.apply(spec, *xs, *weights, *biases)


# ------------------------------------------------------------------------------------------------
# log_softmax, nll, cross entropy
# ------------------------------------------------------------------------------------------------
class _LogSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        L.require_cuda(x, 'log_softmax input')
        x = x.contiguous()
        out = torch.empty_like(x)
        check(lib().agx_log_softmax_nll(ptr(x), x.stride(0), x.shape[0], x.shape[1], None, None,
                                        ptr(out), out.stride(0), None, None, stream_ptr()),
              'agx_log_softmax_nll')
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, g):
        (logp,) = ctx.saved_tensors
        g = g.contiguous()
        dx = torch.empty_like(logp)
        check(lib().agx_log_softmax_nll_bwd(ptr(logp), logp.stride(0), logp.shape[0], logp.shape[1],
                                            None, None, None, None, 1.0, ptr(g), g.stride(0),
                                            ptr(dx), dx.stride(0), stream_ptr()),
              'agx_log_softmax_nll_bwd')
        return dx


class _LogSoftmaxManyFn(torch.autograd.Function):
    """log_softmax(dim=1) of several [N_t, C] tensors that are adjacent row blocks of one buffer
    (the outputs of one hetero conv layer): ONE launch over the whole buffer."""

    @staticmethod
    def forward(ctx, *xs):
        ctx.set_materialize_grads(False)
        base, rows, c = xs[0], sum(x.shape[0] for x in xs), xs[0].shape[1]
        out = torch.empty(rows, c, dtype=torch.float32, device=base.device)
        check(lib().agx_log_softmax_nll(ptr(base), c, rows, c, None, None, ptr(out), c, None, None,
                                        stream_ptr()), 'agx_log_softmax_nll')
        outs, r0 = [], 0
        for x in xs:
            outs.append(out[r0:r0 + x.shape[0]])
            r0 += x.shape[0]
        ctx.save_for_backward(*outs)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        res = []
        for logp, g in zip(ctx.saved_tensors, gs):
            if g is None:
                res.append(None)
                continue
            g = g.contiguous()
            dx = torch.empty(logp.shape, dtype=torch.float32, device=logp.device)
            check(lib().agx_log_softmax_nll_bwd(ptr(logp), logp.stride(0), logp.shape[0],
                                                logp.shape[1], None, None, None, None, 1.0, ptr(g),
                                                g.stride(0), ptr(dx), dx.stride(0), stream_ptr()),
                  'agx_log_softmax_nll_bwd')
            res.append(dx)
        return tuple(res)


def log_softmax_many(xs: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """``[F.log_softmax(x, dim=1) for x in xs]``; one launch when the inputs are consecutive row
    blocks of one contiguous float32 buffer (as _HeteroConvFn lays its outputs out)."""
    xs = list(xs)
    ok = len(xs) > 1 and all(x.dim() == 2 and x.is_cuda and x.dtype == torch.float32 and
                             x.is_contiguous() and x.shape[1] == xs[0].shape[1] for x in xs)
    if ok:
        p = xs[0].data_ptr()
        for x in xs:
            if x.data_ptr() != p:
                ok = False
                break
            p += x.numel() * 4
    if not ok:
        return [log_softmax(x, 1) for x in xs]
    res = list(_LogSoftmaxManyFn.apply(*xs))
    for r, x in zip(res, xs):
        # nll_loss on exactly this tensor differentiates log_softmax + nll in one kernel (below)
        r._agx_logits = x
    return res


def log_softmax(x: torch.Tensor, dim: int = 1) -> torch.Tensor:
    if x.dim() != 2 or dim not in (1, -1):
        raise NotImplementedError('log_softmax: only 2-D input, dim=1')
    r = _LogSoftmaxFn.apply(x)
    r._agx_logits = x          # see log_softmax_many
    return r


class _SoftmaxNLLFn(torch.autograd.Function):
    """coef * weighted-mean nll(log_softmax(logits), labels); also returns log-probabilities."""

    @staticmethod
    def forward(ctx, logits, labels, class_w, coef, group=None):
        L.require_cuda(logits, 'logits')
        logits = logits.contiguous()
        n, c = logits.shape
        dev = logits.device
        labels = labels.to(device=dev, dtype=torch.int64).contiguous()
        logp = torch.empty_like(logits)
        loss_sum = torch.empty(2, dtype=torch.float32, device=dev)
        row_ws = torch.empty(2 * max(n, 1), dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        check(lib().agx_log_softmax_nll(ptr(logits), logits.stride(0), n, c, ptr(labels),
                                        ptr(class_w), ptr(logp), logp.stride(0), ptr(loss_sum),
                                        ptr(row_ws), stream_ptr()), 'agx_log_softmax_nll')
        if group is not None:
            import torch.distributed as dist
            from .dist import small_all_reduce_
            small_all_reduce_(loss_sum, group)
        check(lib().agx_loss_finish(ptr(loss_sum), coef, ptr(loss), 0, stream_ptr()),
              'agx_loss_finish')
        ctx.save_for_backward(logp, labels, loss_sum, class_w if class_w is not None else logp)
        ctx.has_w = class_w is not None
        ctx.coef = coef
        ctx.mark_non_differentiable(logp)
        return loss.reshape(()), logp

    @staticmethod
    def backward(ctx, gloss, _glogp):
        logp, labels, loss_sum, cw = ctx.saved_tensors
        gs = gloss.reshape(1).to(torch.float32).contiguous()
        dx = torch.empty_like(logp)
        check(lib().agx_log_softmax_nll_bwd(ptr(logp), logp.stride(0), logp.shape[0], logp.shape[1],
                                            ptr(labels), ptr(cw) if ctx.has_w else None,
                                            ptr(loss_sum), ptr(gs), ctx.coef, None, 0, ptr(dx),
                                            dx.stride(0), stream_ptr()), 'agx_log_softmax_nll_bwd')
        return dx, None, None, None, None


def cross_entropy(logits, labels, weight: Optional[torch.Tensor] = None, coef: float = 1.0,
                  group=None):
    """``coef * F.cross_entropy(logits, labels, weight)`` (weighted mean), fused fwd/bwd.  With
    ``group`` the weighted mean is taken over the batch shards of all ranks."""
    return _SoftmaxNLLFn.apply(logits, labels, weight, coef, group)[0]


def softmax_nll(logits, labels, weight=None, coef: float = 1.0, group=None):
    """Returns (loss, log_softmax(logits)) -- the GNN's output + loss in one pass."""
    return _SoftmaxNLLFn.apply(logits, labels, weight, coef, group)


class _NLLFromLogpFn(torch.autograd.Function):
    """F.nll_loss on log-probabilities already computed by log_softmax (reference call shape:
    src/train_gnn_embeddings.py:29-30)."""

    @staticmethod
    def forward(ctx, logp, labels, group=None, global_count=None, pending=None):
        L.require_cuda(logp, 'nll_loss input')
        n, c = logp.shape
        dev = logp.device
        labels = labels.to(device=dev, dtype=torch.int64).contiguous()
        lp = logp.contiguous()
        loss_sum = torch.empty(2, dtype=torch.float32, device=dev)
        row_ws = torch.empty(2 * max(n, 1), dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        check(lib().agx_nll_forward(ptr(lp), lp.stride(0), n, c, ptr(labels), None, ptr(loss_sum),
                                    ptr(row_ws), stream_ptr()), 'agx_nll_forward')
        ctx.coef = 1.0
        if group is not None and global_count is not None:
            # mean over the rows of ALL ranks, their number known beforehand: this rank's term is
            # sum_local / N_global -- the backward pass needs nothing from the other ranks, and the
            # all-reduce of the loss VALUE leaves the critical path (async; the caller waits on
            # ``pending`` when it reads the loss)
            import torch.distributed as dist
            ctx.coef = float(n) / float(global_count)
            check(lib().agx_loss_finish(ptr(loss_sum), ctx.coef, ptr(loss), 0, stream_ptr()),
                  'agx_loss_finish')
            from .dist import _PEER, small_all_reduce_
            import os
            if _PEER.get(id(group)) and os.environ.get('AGX_PEER_LOSS', '1') != '0':
                small_all_reduce_(loss, group)          # one small kernel on this stream
            else:
                work = dist.all_reduce(loss, group=group, async_op=True)
                if pending is not None:
                    pending.append(work)
                else:
                    work.wait()
        else:
            if group is not None:    # (sum nll, count) over the rows of all ranks
                from .dist import small_all_reduce_
                small_all_reduce_(loss_sum, group)
            check(lib().agx_loss_finish(ptr(loss_sum), 1.0, ptr(loss), 0, stream_ptr()),
                  'agx_loss_finish')
        ctx.to_save = (labels, loss_sum)
        ctx.save_for_backward(labels, loss_sum)
        ctx.shape = (n, c)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        labels, loss_sum = ctx.saved_tensors
        n, c = ctx.shape
        dlp = torch.empty(n, c, dtype=torch.float32, device=labels.device)
        gs = g.reshape(1).to(torch.float32).contiguous()
        check(lib().agx_nll_backward(n, c, ptr(labels), None, ptr(loss_sum), ptr(gs), ctx.coef,
                                     ptr(dlp), c, stream_ptr()), 'agx_nll_backward')
        return dlp, None, None, None, None


class _NLLOfLogSoftmaxFn(torch.autograd.Function):
    """nll_loss(log_softmax(logits)) when the log-probabilities are the tensor log_softmax_many
    returned: same forward arithmetic as _NLLFromLogpFn (on the log-probabilities), but the
    gradient goes to the LOGITS in one kernel -- g * (softmax - onehot) -- instead of a dense
    [N, C] d(logp) written by one kernel and folded through log_softmax's backward by another."""

    @staticmethod
    def forward(ctx, logits, logp, labels, group, global_count, pending):
        loss = _NLLFromLogpFn.forward(ctx, logp, labels, group, global_count, pending)
        labels_, loss_sum = ctx.to_save
        ctx.save_for_backward(labels_, loss_sum, logp)
        return loss

    @staticmethod
    def backward(ctx, g):
        labels, loss_sum, logp = ctx.saved_tensors
        n, c = ctx.shape
        gs = g.reshape(1).to(torch.float32).contiguous()
        dx = torch.empty(n, c, dtype=torch.float32, device=logp.device)
        check(lib().agx_log_softmax_nll_bwd(ptr(logp), logp.stride(0), n, c, ptr(labels), None,
                                            ptr(loss_sum), ptr(gs), ctx.coef, None, 0, ptr(dx), c,
                                            stream_ptr()), 'agx_log_softmax_nll_bwd')
        return dx, None, None, None, None, None


def nll_loss(logp: torch.Tensor, labels: torch.Tensor, group=None, global_count=None,
             pending=None) -> torch.Tensor:
    """``F.nll_loss(logp, labels)``; with ``group`` the mean runs over the rows of all ranks (each
    rank then holds the gradient of the GLOBAL loss w.r.t. its own rows).  ``global_count``: the
    number of rows over all ranks when known beforehand -- the collective then only carries the
    loss value and runs asynchronously (handles appended to ``pending``)."""
    src = getattr(logp, '_agx_logits', None)
    if src is not None and torch.is_grad_enabled() and src.requires_grad and \
            src.shape == logp.shape and logp.is_contiguous():
        return _NLLOfLogSoftmaxFn.apply(src, logp.detach(), labels, group, global_count, pending)
    return _NLLFromLogpFn.apply(logp, labels, group, global_count, pending)


# ------------------------------------------------------------------------------------------------
# heads: concat-free (feat | emb) -> dropout -> Linear ; projector Linear ; SmoothL1
# ------------------------------------------------------------------------------------------------
class _FusedLinearFn(torch.autograd.Function):
    """out = (cat(parts, dim=1) * cat(masks, dim=1)) @ W^T + b without materialising the
    concatenation: every part is one K-segment of the GEMM against a column slice of W, its
    dropout mask (same shape as the part, nullable) is applied while the tile is staged."""

    @staticmethod
    def forward(ctx, weight, bias, n_parts, *parts_and_masks):
        parts = tuple(p.contiguous() for p in parts_and_masks[:n_parts])
        masks = tuple(parts_and_masks[n_parts:])
        dev = weight.device
        B = parts[0].shape[0]
        Cn = weight.shape[0]
        out = torch.empty(B, Cn, dtype=torch.float32, device=dev)
        segs = []
        off = 0
        for p, m in zip(parts, masks):
            L.require_cuda(p, 'head input')
            if p.dtype != torch.float32:
                raise TypeError('head inputs must be float32')
            k = p.shape[1]
            if m is not None and (m.shape != p.shape or m.stride() != p.stride()):
                raise ValueError('dropout mask must have the shape and strides of its input part')
            segs.append((p, _t(weight[:, off:off + k]), m, None))
            off += k
        if off != weight.shape[1]:
            raise ValueError(f'head inputs have {off} columns, weight expects {weight.shape[1]}')
        gb = ops.GemmBatch()
        gb.add(out, segs, bias=bias)
        gb.run()
        ctx.n_parts = n_parts
        ctx.mask_present = [m is not None for m in masks]
        ctx.has_bias = bias is not None
        ctx.save_for_backward(weight, *parts, *[m for m in masks if m is not None])
        return out

    @staticmethod
    def backward(ctx, g):
        sv = ctx.saved_tensors
        n = ctx.n_parts
        weight, parts = sv[0], sv[1:1 + n]
        it = iter(sv[1 + n:])
        masks = [next(it) if present else None for present in ctx.mask_present]
        g = g.contiguous()
        dev = g.device
        B, Cn = g.shape
        dW = torch.empty_like(weight)
        gb = ops.GemmBatch()
        off = 0
        dparts: List[Optional[torch.Tensor]] = []
        sk = ops.split_k_for(B)
        for i, (p, m) in enumerate(zip(parts, masks)):
            k = p.shape[1]
            gb.add(dW[:, off:off + k], [(_t(g), p, None, m)], split_k=sk)   # g^T @ (p * mask)
            if ctx.needs_input_grad[3 + i]:
                dp = torch.empty_like(p)
                gb.add(dp, [(g, weight[:, off:off + k])])
                dparts.append(dp)
            else:
                dparts.append(None)
            off += k
        gb.run()
        outs = [None if d is None else (ops.scale_mask(d, m) if m is not None else d)
                for d, m in zip(dparts, masks)]
        db = None
        if ctx.has_bias:
            db = torch.empty(Cn, dtype=torch.float32, device=dev)
            ops.colsum([(g, db, False)])
        return (dW, db, None, *outs, *([None] * n))


def fused_linear(parts: Sequence[torch.Tensor], weight, bias=None,
                 dmasks: Optional[Sequence[Optional[torch.Tensor]]] = None):
    if dmasks is None:
        dmasks = [None] * len(parts)
    return _FusedLinearFn.apply(weight, bias, len(parts), *parts, *dmasks)


class _MaskFn(torch.autograd.Function):
    """y = x * mask (dropout with a precomputed multiplicative mask)."""

    @staticmethod
    def forward(ctx, x, mask):
        ctx.save_for_backward(mask)
        return ops.scale_mask(x.contiguous(), mask)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return ops.scale_mask(g.contiguous(), mask), None


def dropout_apply(x, mask):
    return _MaskFn.apply(x, mask)


class _SmoothL1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, target):
        L.require_cuda(out, 'smooth_l1 input')
        out = out.contiguous()
        target = target.to(out.device).contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=out.device)
        dout = torch.empty_like(out)
        ws = torch.empty(lib().agx_smooth_l1_workspace_floats(), dtype=torch.float32,
                         device=out.device)
        check(lib().agx_smooth_l1(ptr(out), ptr(target), out.numel(), ptr(loss), ptr(dout), ptr(ws),
                                  stream_ptr()), 'agx_smooth_l1')
        ctx.save_for_backward(dout)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dout,) = ctx.saved_tensors
        return dout * g, None


def smooth_l1_loss(out, target):
    """``torch.nn.SmoothL1Loss()`` (beta=1, mean), src/train_projector.py:33,52."""
    return _SmoothL1Fn.apply(out, target)


class _MSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, target):
        L.require_cuda(out, 'mse input')
        out = out.contiguous()
        target = target.to(out.device).contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=out.device)
        dout = torch.empty_like(out)
        ws = torch.empty(lib().agx_smooth_l1_workspace_floats(), dtype=torch.float32,
                         device=out.device)
        check(lib().agx_mse(ptr(out), ptr(target), out.numel(), ptr(loss), ptr(dout), ptr(ws),
                            stream_ptr()), 'agx_mse')
        ctx.save_for_backward(dout)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dout,) = ctx.saved_tensors
        return dout * g, None


def mse_loss(out, target):
    """``torch.nn.MSELoss()`` (mean), the Castellano encoder criterion
    (src/train_baseline_context.py:51-53)."""
    return _MSEFn.apply(out, target)


class _TanhFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        L.require_cuda(x, 'tanh input')
        x = x.contiguous()
        y = torch.empty_like(x)
        check(lib().agx_tanh(ptr(x), ptr(y), x.numel(), stream_ptr()), 'agx_tanh')
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        g = g.contiguous()
        dx = torch.empty_like(y)
        check(lib().agx_tanh_bwd(ptr(y), ptr(g), ptr(dx), y.numel(), stream_ptr()), 'agx_tanh_bwd')
        return dx


def tanh(x):
    return _TanhFn.apply(x)


# ------------------------------------------------------------------------------------------------
# GATConv: attention layer over the self-loop augmented relations
# ------------------------------------------------------------------------------------------------
class GATPlan:
    """CSR (by destination) and CSC (by source) of one relation's edge list as GATConv sees it:
    self loops removed, then (i, i) for i < min(N_src, N_dst) appended (PyG 2.0.2, also for
    bipartite edge types).  Index preparation only; cached per edge_index tensor.

    ``csr_row[e]``  destination of CSR slot e (the row index agx_sddmm needs per slot);
    ``csc2csr[q]``  CSR slot of the edge in CSC slot q (per-edge arrays live in CSR order);
    ``csc_pos``     the CSC with ``col = csc2csr``: gathers per-edge scalars by source row."""

    _cache: "dict" = {}

    def __init__(self, edge_index: torch.Tensor, n_src: int, n_dst: int, add_self_loops: bool):
        L.require_cuda(edge_index, 'edge_index')
        ei = edge_index
        if add_self_loops:
            keep = ei[0] != ei[1]
            loops = torch.arange(min(n_src, n_dst), dtype=ei.dtype, device=ei.device)
            ei = torch.cat([ei[:, keep], torch.stack([loops, loops])], dim=1)
        self.n_src, self.n_dst, self.n_edges = int(n_src), int(n_dst), int(ei.shape[1])
        self.edge_index = ei.contiguous()
        self.csr, self.csc = ops.csr_build([(self.edge_index[1], self.edge_index[0], n_dst, n_src),
                                            (self.edge_index[0], self.edge_index[1], n_src, n_dst)])
        E = self.n_edges
        dev = ei.device
        for c in (self.csr, self.csc):
            c.max_degree = int((c.rowptr[1:] - c.rowptr[:-1]).max()) if c.n_rows > 0 and E > 0 else 0
        csr_eid = self.csr.eid[:E].long()
        self.csr_row = self.edge_index[1][csr_eid].to(torch.int32).contiguous()
        inv = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
        inv[csr_eid] = torch.arange(E, dtype=torch.int32, device=dev)
        self.csc2csr = inv[self.csc.eid[:E].long()].contiguous()
        import dataclasses
        self.csc_pos = dataclasses.replace(self.csc, col=self.csc2csr, n_cols=max(E, 1))
        self.long_rows = ops.gat_long_rows(self.csr)       # hub destinations: a CTA each

    @classmethod
    def get(cls, edge_index, n_src, n_dst, add_self_loops=True):
        key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, int(n_src),
               int(n_dst), bool(add_self_loops))
        plan = cls._cache.get(key)
        if plan is None:
            if len(cls._cache) > 64:
                cls._cache.clear()
            plan = cls._cache[key] = cls(edge_index, n_src, n_dst, add_self_loops)
            plan._keepalive = edge_index
        return plan


@dataclass
class GATRelSpec:
    plan: GATPlan
    src: str                      # node-type keys (into GATSpec.node_types)
    dst: str
    i_wl: int                     # indices into the flat parameter list
    i_wr: int
    i_al: int
    i_ar: int
    i_b: int                      # -1: no bias


@dataclass
class GATSpec:
    """Static description of one attention layer (all relations) for _HeteroGATFn."""
    node_types: List[str]
    rels: List[GATRelSpec]
    out_channels: int
    slope: float
    identity: Dict[str, bool] = field(default_factory=dict)   # type -> x[type] is eye(N)
    param_refs: Optional[list] = None

    @property
    def dst_types(self) -> List[str]:
        out: List[str] = []
        for rs in self.rels:
            if rs.dst not in out:
                out.append(rs.dst)
        return out


def _gat_logit_rows(spec: "GATSpec") -> "Dict[str, list]":
    """Per node type the (relation index, role) pairs whose attention logits are computed from that
    type's features: role 'l' = the type is the relation's source, 'r' = its destination."""
    slots: Dict[str, list] = {}
    for k, rs in enumerate(spec.rels):
        slots.setdefault(rs.src, []).append((k, 'l'))
        slots.setdefault(rs.dst, []).append((k, 'r'))
    return slots


def _sum_many(items: list):
    """(out, inputs, bias) sums that do not depend on one another: ONE launch for those with at
    most 8 inputs, the (rare) longer ones chained."""
    if not items:
        return
    ops.sum_arrays([it for it in items if len(it[1]) <= 8])
    for out, ins, bias in items:
        if len(ins) > 8:
            _chain_sum(out, ins, bias)


def _chain_sum(out: torch.Tensor, ins: list, bias: Optional[torch.Tensor] = None):
    """out = sum(ins) (+ bias rows), at most 8 inputs per descriptor (chained in place)."""
    while len(ins) > 8:
        ops.sum_arrays([(out, ins[:8])])
        ins = [out] + ins[8:]
    ops.sum_arrays([(out, ins, bias)])


class _HeteroGATFn(torch.autograd.Function):
    """All GATConv relations of one hetero layer (PyG 2.0.2 GATConv inside to_hetero, aggr='sum'):

        x_l = x_src W_l^T;  a_l = x_src (W_l^T att_l);  a_r = x_dst (W_r^T att_r)
        alpha_ij = softmax_i(leaky_relu(a_l[j] + a_r[i]));  out[t] = sum_{r: dst(r)=t} (sum_j alpha_ij x_l[j] + b_r)

    The dense parts run as two grouped-GEMM waves: the logits are matrix-vector products with the
    projected attention vectors (PyG forms x_l att_l and (x_dst W_r^T) att_r; same value, and
    x_dst W_r^T [N_dst, C] is never needed), stacked per node type so that every feature table is
    read once.  The attention softmax of every relation is ONE scalar launch, the weighted
    neighbour sums of every relation run on the edge-balanced aggregation kernels with per-edge
    weights."""

    @staticmethod
    def forward(ctx, spec: GATSpec, *tensors):
        ctx.set_materialize_grads(False)
        ctx.anchor = bool(getattr(spec, 'anchor', False))
        if ctx.anchor:                        # trailing dummy input, see _direct_call
            tensors = tensors[:-1]
        nt = len(spec.node_types)
        xs = {t: tensors[i] for i, t in enumerate(spec.node_types)}
        params = tensors[nt:]
        for t, x in xs.items():
            L.require_cuda(x, f'x[{t}]')
            if x.dtype != torch.float32 or not x.is_contiguous():
                raise TypeError(f'x[{t}] must be contiguous float32')
        C_ = spec.out_channels
        dev = tensors[0].device
        f32 = dict(dtype=torch.float32, device=dev)

        # f1: x_l = x_src W_l^T (one-hot sources: W_l^T) and, per node type, the stacked logit
        # projections UV[t]: row (k, 'l') = att_l_k W_l_k for the relations t feeds, row (k, 'r') =
        # att_r_k W_r_k for the relations it receives -- a_l = x_src (W_l^T att_l), a_r likewise
        slots = _gat_logit_rows(spec)
        UV = {t: torch.empty(len(rows), xs[t].shape[1], **f32) for t, rows in slots.items()}
        gb = ops.GemmBatch()
        tr: list = []
        x_l = []
        for k, rs in enumerate(spec.rels):
            p = rs.plan
            y = torch.empty(p.n_src, C_, **f32)
            x_l.append(y)
            if spec.identity.get(rs.src, False):
                tr.append((y, params[rs.i_wl]))
            else:
                gb.add(y, [(xs[rs.src], _t(params[rs.i_wl]))])
        for t, rows in slots.items():
            for i, (k, role) in enumerate(rows):
                rs = spec.rels[k]
                att, w = (rs.i_al, rs.i_wl) if role == 'l' else (rs.i_ar, rs.i_wr)
                gb.add(UV[t][i:i + 1], [(params[att].view(1, C_), params[w])])
        if tr:
            ops.transpose_many(tr)
        gb.run()
        # f2: all logits of a node type in ONE product A[t] = UV[t] x[t]^T (x[t] is read once, not
        # once per relation); one-hot features: A[t] = UV[t]
        gb = ops.GemmBatch()
        A = {}
        for t, rows in slots.items():
            if spec.identity.get(t, False):
                A[t] = UV[t]
            else:
                A[t] = torch.empty(len(rows), xs[t].shape[0], **f32)
                gb.add(A[t], [(UV[t], _t(xs[t]))])
        gb.run()
        a_l = [A[rs.src][slots[rs.src].index((k, 'l'))] for k, rs in enumerate(spec.rels)]
        a_r = [A[rs.dst][slots[rs.dst].index((k, 'r'))] for k, rs in enumerate(spec.rels)]
        # f3: attention coefficients of every relation, CSR order
        alpha = [torch.empty(max(rs.plan.n_edges, 1), **f32) for rs in spec.rels]
        ops.gat_edge_softmax([ops.GatArg(rs.plan.csr, a_l[k], a_r[k], alpha[k],
                                         long_rows=rs.plan.long_rows)
                              for k, rs in enumerate(spec.rels)], spec.slope)
        # f4: weighted neighbour sums, relation sum and biases per destination type
        outs: Dict[str, torch.Tensor] = {}
        rows_waves: Dict[int, list] = {}
        chunks: list = []
        finals: list = []
        bias_sums: list = []
        # all destination types' outputs are row blocks of ONE buffer (consumers that run the same
        # row-wise op on every type -- log_softmax -- then need a single launch)
        n_of = {t: next(rs.plan.n_dst for rs in spec.rels if rs.dst == t) for t in spec.dst_types}
        out_all = torch.empty(sum(n_of.values()), C_, **f32)
        row0 = 0
        for t in spec.dst_types:
            lst = [(k, rs) for k, rs in enumerate(spec.rels) if rs.dst == t]
            n_t = n_of[t]
            biases = [params[rs.i_b] for _, rs in lst if rs.i_b >= 0]
            if not biases:
                bias = None
            elif len(biases) == 1:
                bias = biases[0]
            else:
                bias = torch.empty(C_, **f32)
                bias_sums.append((bias, biases))
            short = [(k, rs) for k, rs in lst if not rs.plan.csr.long_rows]
            long_ = [(k, rs) for k, rs in lst if rs.plan.csr.long_rows]
            out = out_all[row0:row0 + n_t]
            row0 += n_t
            for base in range(0, len(short), L.MAX_REL_PER_GROUP):
                part = short[base:base + L.MAX_REL_PER_GROUP]
                rows_waves.setdefault(base // L.MAX_REL_PER_GROUP, []).append(
                    (out, [ops.RelArg(rs.plan.csr, x_l[k], edge_w=alpha[k]) for k, rs in part],
                     base > 0, bias if base == 0 else None))
            temps = []
            for k, rs in long_:
                tmp = out if (not short and len(long_) == 1 and bias is None) else \
                    torch.empty(n_t, C_, **f32)
                temps.append(tmp)
                chunks.append((tmp, ops.RelArg(rs.plan.csr, x_l[k], edge_w=alpha[k])))
            if temps and temps[0] is not out:
                finals.append((out, ([out] if short else []) + temps, None if short else bias))
            outs[t] = out
        _sum_many([(b, list(items), None) for b, items in bias_sums])
        for wave in sorted(rows_waves):
            ops.aggregate_rows(rows_waves[wave], C_)
        ops.aggregate_chunks(chunks, C_)
        _sum_many(finals)

        ctx.spec, ctx.nt = spec, nt
        ctx.slot_types = list(slots.keys())
        ctx.save_for_backward(*tensors, *x_l, *alpha, *[UV[t] for t in ctx.slot_types],
                              *[A[t] for t in ctx.slot_types])
        return tuple(outs[t] for t in spec.dst_types)

    @staticmethod
    def backward(ctx, *douts):
        spec: GATSpec = ctx.spec
        nt = ctx.nt
        saved = ctx.saved_tensors
        R = len(spec.rels)
        T = len(ctx.slot_types)
        n_in = len(saved) - 2 * R - 2 * T
        tensors = saved[:n_in]
        x_l, alpha = saved[n_in:n_in + R], saved[n_in + R:n_in + 2 * R]
        UV = dict(zip(ctx.slot_types, saved[n_in + 2 * R:n_in + 2 * R + T]))
        A = dict(zip(ctx.slot_types, saved[n_in + 2 * R + T:]))
        slots = _gat_logit_rows(spec)
        a_l = [A[rs.src][slots[rs.src].index((k, 'l'))] for k, rs in enumerate(spec.rels)]
        a_r = [A[rs.dst][slots[rs.dst].index((k, 'r'))] for k, rs in enumerate(spec.rels)]
        xs = {t: tensors[i] for i, t in enumerate(spec.node_types)}
        params = tensors[nt:]
        C_ = spec.out_channels
        dev = tensors[0].device
        f32 = dict(dtype=torch.float32, device=dev)
        dout: Dict[str, Optional[torch.Tensor]] = {}
        for t, g in zip(spec.dst_types, douts):
            dout[t] = None if g is None else g.contiguous()
        need_x = {t: ctx.needs_input_grad[1 + i] for i, t in enumerate(spec.node_types)}
        grads: List[Optional[torch.Tensor]] = [None] * len(tensors)
        pidx = lambda i: nt + i                                            # noqa: E731
        live = [(k, rs) for k, rs in enumerate(spec.rels) if dout[rs.dst] is not None]
        if not live:
            return (None, *grads)

        # b1: bias gradients = column sums of dout, shared by the relations of a type
        dbias: Dict[str, torch.Tensor] = {}
        cs_items = []
        for t in spec.dst_types:
            if dout[t] is not None and any(rs.i_b >= 0 for _, rs in live if rs.dst == t):
                dbias[t] = torch.empty(C_, **f32)
                cs_items.append((dout[t], dbias[t], False))
        ops.colsum(cs_items)
        for k, rs in live:
            if rs.i_b >= 0:
                grads[pidx(rs.i_b)] = dbias[rs.dst]

        # b2: d alpha_ij = <dout[i], x_l[j]> per CSR slot;  b3: softmax / leaky-relu chain
        dalpha = {k: torch.empty(max(rs.plan.n_edges, 1), **f32) for k, rs in live}
        ops.sddmm([(rs.plan.csr_row, rs.plan.csr.col, dout[rs.dst], x_l[k], dalpha[k])
                   for k, rs in live if rs.plan.n_edges > 0], C_)
        de = {k: torch.empty(max(rs.plan.n_edges, 1), **f32) for k, rs in live}
        # logit gradients land in the rows of dA[t] (the layout of A[t]); rows of relations whose
        # output received no gradient are cleared
        live_k = {k for k, _ in live}
        dA = {}
        shapes = {t: (len(rows), xs[t].shape[0]) for t, rows in slots.items()
                  if any(k in live_k for k, _ in rows)}
        flat = torch.empty(sum(a * b for a, b in shapes.values()), **f32)
        if any(k not in live_k for t in shapes for k, _ in slots[t]):
            ops.fill_(flat, 0.0)
        off = 0
        for t, (a, b) in shapes.items():
            dA[t] = flat[off:off + a * b].view(a, b)
            off += a * b
        da_l = {k: dA[rs.src][slots[rs.src].index((k, 'l'))] for k, rs in live}
        da_r = {k: dA[rs.dst][slots[rs.dst].index((k, 'r'))] for k, rs in live}
        ops.gat_edge_softmax([ops.GatArg(rs.plan.csr, a_l[k], a_r[k], alpha[k], dalpha=dalpha[k],
                                         de=de[k], da_r=da_r[k], long_rows=rs.plan.long_rows)
                              for k, rs in live], spec.slope, backward=True)

        # b4: transposes over the CSC: dX_l = sum_i alpha_ij dout[i], da_l[j] = sum_i de_ij
        dxl = {k: torch.empty(rs.plan.n_src, C_, **f32) for k, rs in live}
        rows_w, chunks_w, rows_1, chunks_1 = [], [], [], []
        for k, rs in live:
            p = rs.plan
            wide = ops.RelArg(p.csc, dout[rs.dst], edge_w=alpha[k], edge_w_idx=p.csc2csr)
            one = ops.RelArg(p.csc_pos, de[k].view(-1, 1))
            if p.csc.long_rows:
                chunks_w.append((dxl[k], wide))
                chunks_1.append((da_l[k].view(-1, 1), one))
            else:
                rows_w.append((dxl[k], [wide], False))
                rows_1.append((da_l[k].view(-1, 1), [one], False))
        ops.aggregate_rows(rows_w, C_)
        ops.aggregate_chunks(chunks_w, C_)
        ops.aggregate_rows(rows_1, 1)
        ops.aggregate_chunks(chunks_1, 1)

        # b5 (wave A): dUV[t] = dA[t] x[t] (one-hot: dA[t]);  dW_l = dX_l^T x_src (one-hot: dX_l^T)
        gb = ops.GemmBatch()
        trb: list = []
        dUV = {}
        # (the weight gradients first: a grouped call takes 24 problems, and the tensor-core
        # split-K problems among them should share one launch)
        for k, rs in live:
            x = xs[rs.src]
            dwl = torch.empty(C_, x.shape[1], **f32)
            grads[pidx(rs.i_wl)] = dwl
            if spec.identity.get(rs.src, False):
                trb.append((dwl, dxl[k]))                        # dX_l^T I
            else:
                gb.add(dwl, [(_t(dxl[k]), x)], split_k=ops.split_k_for(rs.plan.n_src))
        for t in dA:
            if spec.identity.get(t, False):
                dUV[t] = dA[t]
            else:
                dUV[t] = torch.empty_like(UV[t])
                # few output rows, a very long reduction: short slabs (more CTAs in flight)
                gb.add(dUV[t], [(dA[t], xs[t])], split_k=ops.split_k_for(xs[t].shape[0], slab=128))
        if trb:
            ops.transpose_many(trb)
        gb.run()
        # b6 (wave B): u = W_l^T att_l, v = W_r^T att_r  =>  dW += att du^T, d att = W du;
        # dx[t] = sum_{src(k)=t} dX_l W_l + dA[t]^T UV[t]
        gb = ops.GemmBatch()
        for k, rs in live:
            du = dUV[rs.src][slots[rs.src].index((k, 'l'))]
            dv = dUV[rs.dst][slots[rs.dst].index((k, 'r'))]
            fs, fd = du.numel(), dv.numel()
            gb.add(grads[pidx(rs.i_wl)], [(params[rs.i_al].view(C_, 1), du.view(1, fs))],
                   accumulate=True)
            dwr = torch.empty(C_, fd, **f32)
            gb.add(dwr, [(params[rs.i_ar].view(C_, 1), dv.view(1, fd))])
            grads[pidx(rs.i_wr)] = dwr
            datt_l = torch.empty(C_, 1, **f32)
            gb.add(datt_l, [(params[rs.i_wl], du.view(fs, 1))],
                   split_k=ops.split_k_for(fs, slab=64))
            grads[pidx(rs.i_al)] = datt_l
            datt_r = torch.empty(C_, 1, **f32)
            gb.add(datt_r, [(params[rs.i_wr], dv.view(fd, 1))],
                   split_k=ops.split_k_for(fd, slab=64))
            grads[pidx(rs.i_ar)] = datt_r
        for i, t in enumerate(spec.node_types):
            if not need_x[t]:
                continue
            segs = [(dxl[k], params[rs.i_wl]) for k, rs in live if rs.src == t]
            if t in dA:
                segs.append((_t(dA[t]), UV[t]))
            if segs:
                dx = torch.empty_like(xs[t])
                grads[i] = dx
                gb.add(dx, segs)
        gb.run()
        # gradients in the parameters' own shapes (att_l / att_r are [1, 1, C])
        for i in range(nt, len(tensors)):
            if grads[i] is not None and grads[i].shape != tensors[i].shape:
                grads[i] = grads[i].view(tensors[i].shape)
        _deliver_param_grads(spec.param_refs, grads, nt, must=ctx.anchor)
        return (None, *grads) + ((None,) if ctx.anchor else ())


def hetero_gat(spec: GATSpec, x_list: Sequence[torch.Tensor], params: Sequence[torch.Tensor]):
    return _direct_call(_HeteroGATFn, spec, x_list, params)

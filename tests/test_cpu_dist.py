"""Host-side logic of the multi-GPU path (mmac_b200.dist.GraphPartition) on CPU: world_size-2
``gloo`` processes partition the same heterograph by destination node, exchange boundary rows the
way the CUDA path does (pack -> all-gather -> extended table; gradient: reduce-scatter ->
scatter-add; emulated here with plain torch ops, the product uses agx_pack_rows /
agx_unpack_rows_add + NCCL) and must reproduce the oracle's single-process aggregation of the
whole graph: bit-exact forward (same in-row neighbour order), gradients to 1e-6."""
import os
import socket
import sys
import traceback
from collections import OrderedDict

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util  # noqa: F401  (sys.path)
from oracle import graph_oracle as go


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _extend(part, t, x_owned):
    """pack boundary rows -> all-gather -> [owned ; gathered]  (dist._HaloFn.forward on CPU)."""
    B = part.max_boundary[t]
    if B == 0:
        return x_owned
    send = torch.zeros(B, x_owned.shape[1], dtype=x_owned.dtype)
    idx = part.boundary_idx[t].long()
    send[:idx.numel()] = x_owned[idx]
    bufs = [torch.empty_like(send) for _ in range(part.world)]
    dist.all_gather(bufs, send)
    return torch.cat([x_owned] + bufs, dim=0)


def _fold_grad(part, t, d_ext):
    """reduce-scatter of the gathered region + scatter-add (dist._HaloFn.backward on CPU)."""
    n, B = part.n_owned[t], part.max_boundary[t]
    if B == 0:
        return d_ext
    tail = d_ext[n:].clone()
    dist.all_reduce(tail)
    mine = tail[part.rank * B:(part.rank + 1) * B]
    dx = d_ext[:n].clone()
    idx = part.boundary_idx[t].long()
    dx[idx] += mine[:idx.numel()]
    return dx


def _worker(rank, world, port, size, bounds_kind, errq):
    try:
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        dist.init_process_group('gloo', rank=rank, world_size=world)
        import mmac_b200  # noqa: F401
        from mmac_b200 import synth
        from mmac_b200.dist import GraphPartition, split_bounds
        torch.set_num_threads(1)
        if bounds_kind == 'blocks':
            g = synth.replicate(synth.make_artgraph(size, features='dense'), world)
        else:
            g = synth.make_artgraph(size, features='dense')
        ei = go.to_undirected(g.edge_index_dict)
        n = g.num_nodes_dict
        bounds = None
        if bounds_kind == 'uneven':            # an empty shard and a lopsided one
            bounds = {'style': [0, 0, n['style']], 'artwork': [0, n['artwork'] // 5, n['artwork']]}
        part = GraphPartition(ei, n, world, rank, bounds)
        for t in n:
            assert part.bounds[t][0] == 0 and part.bounds[t][-1] == n[t]
            assert part.n_owned[t] == part.bounds[t][rank + 1] - part.bounds[t][rank]
        if bounds_kind == 'blocks':
            assert not part.has_halo and part.halo_rows() == 0
            for t in n:
                assert part.n_ext[t] == part.n_owned[t] == n[t] // world
        else:
            assert part.has_halo

        # every edge lands on exactly one rank
        cnt = torch.tensor([sum(v.shape[1] for v in part.edge_index.values())])
        dist.all_reduce(cnt)
        assert int(cnt) == sum(v.shape[1] for v in ei.values())

        xg = {t: g.x_dict[t].clone().requires_grad_(True) for t in n}
        xl = {t: part.owned(t, g.x_dict[t]).clone().requires_grad_(True) for t in n}
        ext = {t: _extend(part, t, xl[t]) for t in n}
        for t in n:
            eg = part.ext_global[t]
            assert ext[t].shape[0] == part.n_ext[t] == eg.numel()
            ok = eg >= 0
            assert torch.equal(ext[t][ok].detach(), g.x_dict[t][eg[ok]])
        tot_g = 0
        tot_l = 0
        for k, ((s, r, d), e) in enumerate(ei.items()):
            reduce = 'mean' if k % 2 == 0 else 'sum'
            ref = go.propagate(xg[s], e, n[d], reduce)
            le = part.edge_index[(s, r, d)]
            if le.numel():
                assert int(le[0].max()) < part.n_ext[s] and int(le[1].max()) < part.n_owned[d]
            out = go.propagate(ext[s], le, part.n_owned[d], reduce)
            assert torch.equal(out.detach(), part.owned(d, ref).detach()), (s, r, d)
            w = torch.cos(torch.arange(ref.numel(), dtype=torch.float32)).view_as(ref) * (1 + k)
            tot_g = tot_g + (ref * w).sum()
            tot_l = tot_l + (out * part.owned(d, w)).sum()
        tot_g.backward()
        d_ext = torch.autograd.grad(tot_l, [ext[t] for t in n], allow_unused=True)
        for t, de in zip(n, d_ext):
            if de is None:
                de = torch.zeros_like(ext[t])
            dx = _fold_grad(part, t, de)
            ref = part.owned(t, xg[t].grad)
            err = float((dx - ref).abs().max()) / max(float(ref.abs().max()), 1e-30) \
                if ref.numel() else 0.0
            assert err < 1e-6, (t, err)
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errq.put(f'rank {rank}:\n{traceback.format_exc()}')
        raise


@pytest.mark.parametrize('bounds_kind', ['even', 'uneven', 'blocks'])
def test_partition_halo_exchange_world2_gloo(bounds_kind):
    world = 2
    ctx = mp.get_context('spawn')
    errq = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 'tiny', bounds_kind, errq))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    msgs = []
    while not errq.empty():
        msgs.append(errq.get())
    assert not msgs, '\n'.join(msgs)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]


def test_split_bounds():
    import mmac_b200  # noqa: F401
    from mmac_b200.dist import split_bounds
    assert split_bounds(10, 4) == [0, 3, 6, 8, 10]
    assert split_bounds(2, 4) == [0, 1, 2, 2, 2]
    assert split_bounds(0, 2) == [0, 0, 0]
    for n in (1, 7, 18, 32, 116475):
        for w in (1, 2, 4, 8):
            b = split_bounds(n, w)
            sizes = [b[i + 1] - b[i] for i in range(w)]
            assert b[0] == 0 and b[-1] == n and max(sizes) - min(sizes) <= 1


def test_partition_single_rank_is_identity():
    import mmac_b200  # noqa: F401
    from mmac_b200 import synth
    from mmac_b200.dist import GraphPartition
    g = synth.make_artgraph('tiny', features='dense')
    ei = go.to_undirected(g.edge_index_dict)
    part = GraphPartition(ei, g.num_nodes_dict, 1, 0)
    assert not part.has_halo
    for k, v in ei.items():
        assert torch.equal(part.edge_index[k], v)


def test_scattered_partition_index_arithmetic():
    """GraphPartition(scattered=): every edge lands on exactly one rank; relations from artwork into
    a scattered type follow their SOURCE (global destination ids in a table of world * chunk rows,
    divisor = global in-degree); the other relations of a scattered type follow the destination
    and read gathered boundary rows; chunk bounds are equal-sized up to the last one."""
    import mmac_b200  # noqa: F401
    from mmac_b200 import synth
    from mmac_b200.dist import GraphPartition
    g = synth.make_artgraph('small', features='dense')
    ei = go.to_undirected(g.edge_index_dict)
    n = g.num_nodes_dict
    scat = ['tag', 'artist']
    rep = [t for t in n if t != 'artwork' and t not in scat]
    for world in (2, 3, 5):
        parts = [GraphPartition(ei, n, world, r, replicated=rep, scattered=scat) for r in range(world)]
        for t in scat:
            chunk = -(-n[t] // world)
            assert parts[0].chunk[t] == chunk
            assert parts[0].bounds[t] == [min(q * chunk, n[t]) for q in range(world + 1)]
            assert sum(p.n_owned[t] for p in parts) == n[t]
        for et, e in ei.items():
            s_, _, d_ = et
            per_rank = [p.edge_index[et] for p in parts]
            both_rep = s_ in rep and d_ in rep
            total = sum(x.shape[1] for x in per_rank)
            assert total == (world if both_rep else 1) * e.shape[1], et
            if et in parts[0].scatter:
                assert d_ in scat and s_ == 'artwork'
                cnt = torch.bincount(e[1], minlength=world * parts[0].chunk[d_]).clamp(min=1).float()
                for r, p in enumerate(parts):
                    assert torch.equal(p.scatter[et], cnt)
                    lo = p.bounds[s_][r]
                    # sources are owned rows (local index), destinations global ids, edge order kept
                    m = (e[0] >= lo) & (e[0] < p.bounds[s_][r + 1])
                    assert torch.equal(p.edge_index[et][0], e[0][m] - lo)
                    assert torch.equal(p.edge_index[et][1], e[1][m])
            elif d_ in scat:
                for r, p in enumerate(parts):        # destination-owned, local destination ids
                    lo, hi = p.bounds[d_][r], p.bounds[d_][r + 1]
                    m = (e[1] >= lo) & (e[1] < hi)
                    assert torch.equal(p.edge_index[et][1], e[1][m] - lo)
        # tag / artist rows that artworks of another rank read are boundary rows, artwork has none
        assert all(p.max_boundary['artwork'] == 0 for p in parts)
        assert all(p.max_boundary['tag'] > 0 for p in parts)


def test_balanced_bounds_equalise_edges():
    import mmac_b200  # noqa: F401
    from mmac_b200 import synth
    from mmac_b200.dist import GraphPartition, balanced_bounds
    g = synth.make_artgraph('small', features='dense')
    ei = go.to_undirected(g.edge_index_dict)
    n = g.num_nodes_dict
    for world in (2, 4):
        b = balanced_bounds(ei, n, world)
        for t in n:
            assert len(b[t]) == world + 1 and b[t][0] == 0 and b[t][-1] == n[t]
            assert all(b[t][i] <= b[t][i + 1] for i in range(world))
        even = [sum(v.shape[1] for v in GraphPartition(ei, n, world, r).edge_index.values())
                for r in range(world)]
        bal = [sum(v.shape[1] for v in GraphPartition(ei, n, world, r, b).edge_index.values())
               for r in range(world)]
        assert sum(even) == sum(bal) == sum(v.shape[1] for v in ei.values())
        assert max(bal) / min(bal) < max(even) / min(even)
        assert max(bal) <= 1.25 * sum(bal) / world


# ------------------------------------------------------------------------------------------------
# the PRODUCT's multi-GPU path itself (hetero module + dist.HaloExchange + cross-rank BatchNorm /
# loss) on two gloo ranks, every agx entry point restated in torch (tests/cpu_shim.py)
# ------------------------------------------------------------------------------------------------
def _product_worker(rank, world, port, opname, replicate, errq):
    try:
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        dist.init_process_group('gloo', rank=rank, world_size=world)
        torch.set_num_threads(2)
        import mmac_b200 as agx
        from cpu_shim import cpu_ops
        from mmac_b200 import synth
        from mmac_b200.dist import GraphPartition, partition_context
        from mmac_b200.hetero import HeteroModule
        from util import rel_err
        g = synth.make_artgraph('tiny', features='dense')
        ei = go.to_undirected(g.edge_index_dict)
        md = (g.node_types, list(ei.keys()))
        n = g.num_nodes_dict
        y = g['artwork'].y_style
        orc = go.HeteroSGNNOracle(getattr(go, opname), torch.nn.ReLU(), 'sum', 128, 32, md, 2, 0.4,
                                  True, False)
        with torch.no_grad():
            orc(g.x_dict, ei)
        util.fill_params_deterministic(orc)
        util.reset_bn(orc)
        prod = agx.HeteroSGNN(getattr(agx, opname), torch.nn.ReLU(), 'sum', 128, 32, md, 2, 0.4,
                              True, False)
        util.copy_state(orc, prod)
        gen = torch.Generator().manual_seed(77)
        masks = {t: (torch.rand(k, 128, generator=gen) >= 0.4).float() / 0.6 for t, k in n.items()}
        o64 = orc.double().train()
        o64.gnn.dropout_masks = {t: m.double() for t, m in masks.items()}
        e_o, o_o = o64({k: v.double() for k, v in g.x_dict.items()}, ei)
        l_o = go.nll_loss_artwork(o_o[0], y)
        l_o.backward()

        scat = ['tag', 'artist'] if replicate == 'scatter' else []
        rep = [t for t in n if t != 'artwork' and t not in scat] if replicate else []
        part = GraphPartition(ei, n, world, rank, replicated=rep, scattered=scat)
        if replicate == 'scatter':
            # tag / artist rows are owned in equal chunks; artwork -> tag / artist edges stay with
            # their source rank (reduce-scatter), tag -> artwork reads gathered boundary rows
            assert part.has_halo and part.partial and len(part.scatter) == 2
            assert part.max_boundary['artwork'] == 0
        else:
            assert part.has_halo != replicate and (len(part.partial) > 0) == replicate
        ctx = partition_context(part, dist.group.WORLD, 'cpu')
        prod.gnn.dropout_masks = {t: part.owned(t, m).contiguous() for t, m in masks.items()}
        for m in prod.modules():
            if isinstance(m, HeteroModule):
                m.set_distributed(ctx)
        prod.train()
        x_own = {t: part.owned(t, v).contiguous() for t, v in g.x_dict.items()}
        with cpu_ops():
            e_p, o_p = prod(x_own, part.edge_index)
            l_p = agx.functional.nll_loss(o_p[0]['artwork'], part.owned('artwork', y),
                                          dist.group.WORLD)
            l_p.backward()
        assert rel_err(l_p, l_o) <= 1e-5                      # the GLOBAL loss on every rank
        for t in e_o:
            lo, hi = (0, n[t]) if t in part.replicated else \
                (part.bounds[t][rank], part.bounds[t][rank + 1])
            if hi > lo:
                assert rel_err(e_p[t], e_o[t][lo:hi]) <= 2e-5, ('emb', t)
                assert rel_err(o_p[0][t], o_o[0][t][lo:hi]) <= 2e-5, ('logp', t)
        og = {k: p.grad for k, p in o64.named_parameters() if p.grad is not None}
        if opname == 'GraphConv':
            og = {k.replace('.lin_l.', '.lin_rel.').replace('.lin_r.', '.lin_root.'): v
                  for k, v in og.items()}
        pp = dict(prod.named_parameters())
        gmax = max(float(v.abs().max()) for v in og.values())
        for k in sorted(og):
            p = pp[k]
            gk = p.grad.clone() if p.grad is not None else torch.zeros_like(p)
            dist.all_reduce(gk)                               # weight gradients: sum over ranks
            err = float((gk.double() - og[k]).abs().max())
            scale = float(og[k].abs().max())
            # small tensors (biases in front of a training-mode BatchNorm have a zero true gradient,
            # BatchNorm biases are sums of thousands of cancelling terms) carry float32 noise on
            # the scale of the model's gradients, as in tests/test_gpu_model.py
            assert err <= max(1e-4 * scale, 2e-6 * gmax), (k, err, scale)
        for (k, b_o), (_, b_p) in zip(o64.named_buffers(), prod.named_buffers()):
            assert rel_err(b_p, b_o) <= 1e-5, k                # running statistics of ALL rows
        # dropout streams: per rank for the partitioned types, ONE stream for the replicated ones
        hm = prod.gnn
        hm.dropout_masks = None
        with cpu_ops():
            types = list(n.keys())
            ms = hm._masks([(x_own[t].shape[0], 8) for t in types], 0.4, 'cpu', types)
        for t, m in zip(types, ms):
            if m.shape[0] == 0:
                continue
            both = [torch.empty_like(m) for _ in range(world)] if t in part.replicated else None
            if both is not None:
                dist.all_gather(both, m)
                assert all(torch.equal(both[0], b) for b in both[1:]), t
        if not replicate:
            a = ms[0][:8].contiguous()
            both = [torch.empty_like(a) for _ in range(world)]
            dist.all_gather(both, a)
            assert not torch.equal(both[0], both[1])
        # three optimisation steps (gradients now ACCUMULATE into existing .grad buffers, the path
        # the fused layers use with FlatAdam's arena): loss trajectory of the whole-graph oracle
        hm.dropout_masks = {t: part.owned(t, m).contiguous() for t, m in masks.items()}
        o32 = o64.float()
        o32.gnn.dropout_masks = masks
        opt_o = torch.optim.Adam(o32.parameters(), lr=0.01)
        opt_p = torch.optim.Adam(prod.parameters(), lr=0.01)
        for step in range(3):
            opt_o.zero_grad(set_to_none=False)
            _, oo = o32(g.x_dict, ei)
            lo_ = go.nll_loss_artwork(oo[0], y)
            lo_.backward()
            opt_o.step()
            opt_p.zero_grad(set_to_none=False)
            with cpu_ops():
                _, op_ = prod(x_own, part.edge_index)
                lp_ = agx.functional.nll_loss(op_[0]['artwork'], part.owned('artwork', y),
                                              dist.group.WORLD)
                lp_.backward()
            for p_ in prod.parameters():
                if p_.grad is not None:
                    dist.all_reduce(p_.grad)
            opt_p.step()
            # (Adam's first steps are sign-like on noise-level gradients: 2e-3, as on the GPU)
            assert abs(float(lp_) - float(lo_)) <= 2e-3 * abs(float(lo_)), (step, float(lp_), float(lo_))
        # save_embeddings() (train_gnn_embeddings.py:82-93) = deepcopy + eval forward: the cached
        # layer plans hold the process group, which must not be copied (GNNTrainer.embeddings())
        import copy
        clone = copy.deepcopy(prod).eval()
        assert clone.gnn._dist is ctx and not clone.gnn._conv_specs
        with cpu_ops(), torch.no_grad():
            e_c, _ = clone(x_own, part.edge_index)
        prod.eval()
        o32.eval()
        with cpu_ops(), torch.no_grad():
            e_s, _ = prod(x_own, part.edge_index)
            e_r, _ = o32(g.x_dict, ei)
        assert torch.equal(e_c['artwork'], e_s['artwork'])
        lo, hi = part.bounds['artwork'][rank], part.bounds['artwork'][rank + 1]
        # three sign-like Adam steps apart from the oracle's trajectory: a sanity bound only
        assert rel_err(e_c['artwork'], e_r['artwork'][lo:hi]) <= 5e-2
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errq.put(f'rank {rank}:\\n{traceback.format_exc()}')
        raise


@pytest.mark.parametrize('replicate', [False, True])
@pytest.mark.parametrize('opname', ['SAGEConv', 'GraphConv'])
def test_product_cut_partition_world2_gloo(opname, replicate):
    """Two ranks, one graph cut by destination node: embeddings / log-probabilities of the owned
    rows, the global loss, the all-reduced weight gradients and the BatchNorm running statistics
    equal the single-process oracle on the whole graph.  ``replicate``: every node type but
    ``artwork`` lives on both ranks (no boundary rows; partial neighbour sums all-reduced)."""
    _run_product(2, opname, replicate)


@pytest.mark.parametrize('world', [2, 3])
def test_product_scattered_partition_gloo(world):
    """tag and artist cut into equal chunks (their dense work done once, by the owner): relations
    from artwork into them keep their edges with the source rank and reduce-scatter the partial
    sums; with three ranks the last chunk is shorter than the others (padding rows)."""
    _run_product(world, 'SAGEConv', 'scatter')


def test_product_replicated_partition_world4_gloo():
    """Four ranks: the masks / partial sums / gradient parts of the replicated types agree for more
    than two participants."""
    _run_product(4, 'SAGEConv', True)


def _run_product(world, opname, replicate):
    ctx = mp.get_context('spawn')
    errq = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_product_worker, args=(r, world, port, opname, replicate, errq))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    msgs = []
    while not errq.empty():
        msgs.append(errq.get())
    assert not msgs, '\\n'.join(msgs)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]

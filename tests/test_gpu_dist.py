"""Multi-GPU parity (needs >= 2 GPUs on the box: ``gpurun --gpus 2 -- pytest tests -m gpu``): one
process per GPU over NCCL must reproduce the single-GPU step on the whole graph -- embeddings,
log-probabilities, loss and (all-reduced) gradients to rel 1e-5 -- for a destination partition
that cuts edges (boundary-row all-gather / gradient reduce-scatter) and for the block-diagonal
replicated graph (no edge cut); plus the CUDA-graph captured distributed trainer and the
batch-sharded heads."""
import copy
import os
import socket
import traceback
from collections import OrderedDict

import pytest
import torch

import util
from util import rel_err

pytestmark = pytest.mark.gpu
RTOL = 1e-5
GRAD_RTOL = 2e-5


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _model(agx, go, g, ei_cpu, md, opname, C, dev):
    orc = go.HeteroSGNNOracle(getattr(go, opname), torch.nn.ReLU(), 'sum', 128, C, md, 2, 0.0, True,
                              False)
    with torch.no_grad():
        orc(g.x_dict, ei_cpu)
    util.fill_params_deterministic(orc)
    util.reset_bn(orc)

    def make():
        m = agx.HeteroSGNN(getattr(agx, opname), torch.nn.ReLU(), 'sum', 128, C, md, 2, 0.0, True,
                           False)
        util.copy_state(orc, m)
        return m.to(dev).train()
    return orc, make


def _gnn_case(rank, world, dev, kind, opname):
    import torch.distributed as dist
    import mmac_b200 as agx
    from mmac_b200 import synth
    from mmac_b200.dist import GraphPartition, partition_context
    from mmac_b200.trainer import GNNTrainer
    from oracle import graph_oracle as go
    if kind == 'blocks':
        g = synth.replicate(synth.make_artgraph('tiny', features='one-hot'), world)
    else:
        g = synth.make_artgraph('small', features='dense')
    ei_cpu = go.to_undirected(g.edge_index_dict)
    md = (g.node_types, list(ei_cpu.keys()))
    n = g.num_nodes_dict
    C = 32
    y = g['artwork'].y_style.long()
    orc, make = _model(agx, go, g, ei_cpu, md, opname, C, dev)

    # --- single GPU, whole graph ---------------------------------------------------------------
    ref = make()
    xg = OrderedDict((k, v.to(dev)) for k, v in g.x_dict.items())
    eg = OrderedDict((k, v.to(dev)) for k, v in ei_cpu.items())
    emb_r, out_r = ref(xg, eg)
    loss_r = agx.functional.nll_loss(out_r[0]['artwork'], y.to(dev))
    loss_r.backward()
    grads_r = {k: p.grad.clone() for k, p in ref.named_parameters() if p.grad is not None}
    if rank == 0:           # ... which itself matches the CPU oracle
        o2 = copy.deepcopy(orc).train()          # (keeps orc's BatchNorm buffers pristine for make())
        emb_o, out_o = o2(g.x_dict, ei_cpu)
        assert rel_err(emb_r['artwork'], emb_o['artwork']) <= RTOL
        assert rel_err(out_r[0]['artwork'], out_o[0]['artwork']) <= RTOL

    # --- this rank of the partitioned job ----------------------------------------------------------
    # 'replicated': every node type but artwork lives on both ranks -- no boundary
    # rows; the partial neighbour sums of the artwork -> X relations are all-reduced in the layer
    # 'scattered': as 'replicated', but tag and artist are cut into equal chunks -- artwork -> tag /
    # artist partial sums are reduce-scattered to the owners, tag / artist rows all-gathered for the
    # reverse relations
    scat = ['tag', 'artist'] if kind == 'scattered' else []
    part = GraphPartition(eg, n, world, rank, scattered=scat,
                          replicated=[t for t in n if t != 'artwork' and t not in scat]
                          if kind in ('replicated', 'scattered') else ())
    assert part.has_halo == (kind in ('cut', 'scattered'))
    assert (len(part.partial) > 0) == (kind in ('replicated', 'scattered'))
    assert len(part.scatter) == (2 if kind == 'scattered' else 0)
    ctx = partition_context(part, dist.group.WORLD, dev)
    mod = make()
    mod.gnn.set_distributed(ctx)
    xl = OrderedDict((t, part.owned(t, xg[t]).contiguous()) for t in n)
    yl = part.owned('artwork', y).to(dev)
    emb, out = mod(xl, part.edge_index)
    loss = agx.functional.nll_loss(out[0]['artwork'], yl, dist.group.WORLD)
    loss.backward()
    for t in n:
        # BatchNorm over a few dozen rows (style 32, genre 18, field 10 ...) amplifies float32
        # rounding ~100x: the single-GPU float32 reference is itself 2e-6..3e-5 from its float64
        # restatement there; the 1e-5 bound is held on the large types
        tol = RTOL if n[t] >= 512 else 1e-4
        e = rel_err(emb[t], part.owned(t, emb_r[t])) if part.n_owned[t] else 0.0
        assert e <= tol, (kind, 'emb', t, e)
        if part.n_owned[t]:
            # log-probabilities of dead types see no loss but must still be returned and right
            scale = float(out_r[0][t].detach().abs().max())
            d = float((out[0][t] - part.owned(t, out_r[0][t])).detach().abs().max())
            assert d <= tol * scale, (kind, 'logp', t, d / scale)
    assert abs(loss.item() - loss_r.item()) <= RTOL * abs(loss_r.item())
    gmax = max(float(v.abs().max()) for v in grads_r.values())
    for k, p in mod.named_parameters():
        if k not in grads_r:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        gsum = p.grad.clone() if p.grad is not None else torch.zeros_like(grads_r[k])
        dist.all_reduce(gsum)
        # a float32 sum of per-rank partials is exact to the scale of the PARTIALS: bias gradients
        # in front of a training-mode BatchNorm cancel to ~0 only across ranks
        pmax = gsum.new_tensor([float(p.grad.abs().max()) if p.grad is not None else 0.0])
        dist.all_reduce(pmax, op=dist.ReduceOp.MAX)
        d = float((gsum - grads_r[k]).abs().max())
        scale = max(float(grads_r[k].abs().max()), float(pmax))
        # weights of relations that touch a node type of a few dozen rows sit behind a BatchNorm
        # over those rows (rounding amplified ~100x, see above): 1e-4 there, as for the embeddings
        small = any(n[t] < 512 for t in n if f'{t}__' in k or f'__{t}.' in k or f'.{t}.' in k)
        tol = 1e-4 if small else GRAD_RTOL
        util.parity_log('dist-grad', f'{kind} {opname} {k}', d / max(scale, 1e-30), tol=tol)
        assert d <= tol * scale or d <= 1e-5 * gmax, \
            (kind, 'grad', k, d, scale, gmax)
    # BatchNorm running statistics are those of the whole graph
    for (k, b), (_, br) in zip(mod.named_buffers(), ref.named_buffers()):
        if 'running' in k:
            assert rel_err(b, br) <= 1e-5, (kind, opname, k, rel_err(b, br), b[:3].tolist(),
                                            br[:3].tolist())

    # --- trainers: CUDA-graph captured distributed step == single-GPU step -----------------------
    t_ref = GNNTrainer(make(), xg, eg, y, lr=0.01, use_cuda_graph=True)
    t_dist = GNNTrainer(make(), xl, part.edge_index, part.owned('artwork', y), lr=0.01,
                        use_cuda_graph=True, dist_ctx=partition_context(part, dist.group.WORLD, dev))
    for step in range(4):
        lr_ = float(t_ref.train_step().item())
        ld_ = float(t_dist.train_step().item())
        # Adam's first steps are sign-like (m/sqrt(v) ~ +-1) on noise-level gradients, so the
        # trajectories are compared at 2e-3 like the single-GPU optimizer test
        assert abs(lr_ - ld_) <= 2e-3 * abs(lr_), (kind, step, lr_, ld_)
    emb_a = t_ref.embeddings()
    emb_b = t_dist.embeddings()        # deepcopy + eval forward on this rank's partition
    e = rel_err(emb_b['artwork'], part.owned('artwork', emb_a['artwork']))
    assert e <= 2e-2, (kind, 'trained embedding', e)
    torch.cuda.synchronize()


def _heads_case(rank, world, dev):
    """Batch-sharded multitask heads: per-rank shards, class-weighted CE with the all-reduced
    normaliser, all-reduced gradient == single-GPU gradient on the whole batch."""
    import torch.distributed as dist
    import mmac_b200 as agx
    from mmac_b200 import synth
    from mmac_b200.heads import multitask_loss
    Bsz = 256
    feat, es, eg_, ys, yg = [t.to(dev) for t in synth.make_head_batch(Bsz, 'vit')]
    ws = synth.class_weights(ys.cpu(), 32).to(dev)
    wg = synth.class_weights(yg.cpu(), 18).to(dev)
    torch.manual_seed(5)
    ref = agx.NewMultiModalMultiTaskHead(128, {'style': 32, 'genre': 18}, 0.0, 768).to(dev)
    mod = agx.NewMultiModalMultiTaskHead(128, {'style': 32, 'genre': 18}, 0.0, 768).to(dev)
    mod.load_state_dict(ref.state_dict())
    loss_r = multitask_loss(ref(feat, es, eg_), ys, yg, ws, wg)
    loss_r.backward()
    lo, hi = rank * Bsz // world, (rank + 1) * Bsz // world
    sl = slice(lo, hi)
    loss = multitask_loss(mod(feat[sl], es[sl], eg_[sl]), ys[sl], yg[sl], ws, wg,
                          group=dist.group.WORLD)
    loss.backward()
    assert abs(loss.item() - loss_r.item()) <= RTOL * abs(loss_r.item())
    for (k, p), (_, pr) in zip(mod.named_parameters(), ref.named_parameters()):
        gsum = p.grad.clone()
        dist.all_reduce(gsum)
        assert rel_err(gsum, pr.grad) <= GRAD_RTOL, (k, rel_err(gsum, pr.grad))


def _peer_allreduce_case(rank, world, dev):
    """agx_peer_allreduce (one kernel over NVLink peer memory) against NCCL: same sums, bit-identical
    on all ranks, correct over many calls (epoch / slot reuse) and inside a replayed CUDA graph."""
    import torch.distributed as dist
    from mmac_b200 import dist as AD
    ok = AD.enable_peer_allreduce(dist.group.WORLD, dev)
    assert ok, AD.PEER_STATUS
    par = AD._PEER[id(dist.group.WORLD)]
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    for it, (n, dt) in enumerate([(1, torch.float32), (2, torch.float32), (4608, torch.float64),
                                  (45050, torch.float32), (131072, torch.float32), (7, torch.float64)] * 3):
        x = torch.randn(n, generator=gen, device=dev, dtype=dt)
        ref = x.clone()
        dist.all_reduce(ref)
        got = AD.small_all_reduce_(x.clone(), dist.group.WORLD)
        assert torch.allclose(got, ref, rtol=1e-6 if dt == torch.float32 else 1e-13, atol=1e-6), (it, n, dt)
        both = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(both, got)
        assert all(torch.equal(both[0], b) for b in both[1:]), 'ranks disagree bitwise'
    assert par.calls >= 18
    big = torch.ones(1 << 20, device=dev)            # too large for the slot: NCCL path
    assert float(AD.small_all_reduce_(big, dist.group.WORLD)[0]) == world
    # captured: 20 replays of two back-to-back reductions
    a = torch.full((1000,), float(rank + 1), device=dev)
    b = torch.zeros(3, device=dev, dtype=torch.float64)
    out_a, out_b = torch.empty_like(a), torch.empty_like(b)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            out_a.copy_(a); AD.small_all_reduce_(out_a, dist.group.WORLD)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, capture_error_mode='thread_local'):
        out_a.copy_(a); AD.small_all_reduce_(out_a, dist.group.WORLD)
        out_b.copy_(b); AD.small_all_reduce_(out_b, dist.group.WORLD)
    for k in range(20):
        b.fill_(k + rank)
        gr.replay()
        torch.cuda.synchronize()
        assert float(out_a[0]) == world * (world + 1) / 2
        assert float(out_b[0]) == sum(k + r for r in range(world)), (k, out_b)
    dist.barrier()


def _worker(rank, world, port, errq, cases=None):
    try:
        import torch.distributed as dist
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device('cuda', rank)
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
        for kind, op in (cases or (('cut', 'SAGEConv'), ('blocks', 'SAGEConv'), ('cut', 'GraphConv'))):
            _gnn_case(rank, world, dev, kind, op)
        if cases is None:
            _heads_case(rank, world, dev)
            _peer_allreduce_case(rank, world, dev)
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errq.put(f'rank {rank}:\n{traceback.format_exc()}')
        raise


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_ranks_replicated_types_match_single_gpu():
    """Partition with replicated small node types (CPU-verified on gloo in
    tests/test_cpu_dist.py; this is its NCCL / CUDA-graph run)."""
    _run_two_ranks((('replicated', 'SAGEConv'), ('replicated', 'GraphConv')))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_ranks_scattered_types_match_single_gpu():
    """Partition with tag / artist cut into chunks (reduce-scatter of the partial sums into them,
    all-gather of their rows out of them) and the tiny types replicated: the config-5 partition
    of bench.py."""
    _run_two_ranks((('scattered', 'SAGEConv'), ('scattered', 'GraphConv')))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_ranks_match_single_gpu():
    _run_two_ranks(None)


def _run_two_ranks(cases):
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context('spawn')
    errq = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, errq, cases)) for r in range(world)]
    import time
    for p in procs:
        p.start()
    t0 = time.time()
    # a rank that fails leaves its peer waiting inside a collective: stop everything at once
    while any(p.is_alive() for p in procs) and time.time() - t0 < 420 and \
            not any(p.exitcode not in (None, 0) for p in procs):
        time.sleep(0.5)
    msgs = []
    while not errq.empty():
        msgs.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.kill()
            p.join(10)
            msgs.append('worker stopped (peer failed or timed out)')
    assert not msgs, '\n'.join(msgs)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]

#!/usr/bin/env python
"""Benchmark of the hot path: one full-graph hetero-GNN training step (forward + nll_loss +
backward + Adam, the body of ``hetero_training()`` in the reference's train_gnn_embeddings.py)
on a synthetic ArtGraph-shaped heterograph.

    python bench.py --gpus 1 --steps 20 --warmup 5            # this repo's CUDA path
    python bench.py --impl reference --steps 3 --warmup 1     # reference semantics on host CPU cores
    torchrun ... bench.py --gpus N ...                        # one rank per GPU, weak scaling

Prints ONE JSON line (rank 0).  Metric: aggregated edges/s = 5 * sum_r E_r / step time (three
forward aggregation passes + the two transpose passes of the backward; the count is
the reference formulation's and does not depend on how this implementation prunes or reorders).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_JSON_OUT = None
METRIC = 'hetero_gnn_aggregated_edges_per_s'
UNIT = 'edges/s'
PASSES = 5                       # aggregation passes per training step


def _peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader',
                                       '-i', str(self.idx), '-lms', '20'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(',')]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1].split()[0]))
                mx.append(float(c[2].split()[0]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        os.unlink(self.f.name)
        return {'sm_mhz': statistics.median(sm) if sm else None,
                'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# reference semantics on the host CPU (oracle/; test + baseline infrastructure, never the product)
# ------------------------------------------------------------------------------------------------
PARITY_MASK_SEED = 4321


def parity_masks(num_nodes, seed=PARITY_MASK_SEED, p=0.4, hidden=128):
    """Dropout masks injected into BOTH arms of a parity step (the Philox stream of the CUDA path
    cannot reproduce torch's CPU generator)."""
    gen = torch.Generator().manual_seed(seed)
    return OrderedDict((t, (torch.rand(n, hidden, generator=gen) >= p).float() / (1.0 - p))
                       for t, n in num_nodes.items())


def cpu_reference(size: str, steps: int, warmup: int, parity: bool = False):
    """``parity``: the first (warm-up) step runs with injected dropout masks from a recorded initial
    state; its embeddings, log-probabilities, loss and gradients are returned so that the CUDA
    path can be checked against this very run (``gpu_parity``)."""
    from mmac_b200 import synth
    from oracle import graph_oracle as go
    torch.set_num_threads(os.cpu_count() or 1)
    g = synth.make_artgraph(size, features='one-hot')
    ei = go.to_undirected(g.edge_index_dict)
    md = (g.node_types, list(ei.keys()))
    n_edges = sum(int(v.shape[1]) for v in ei.values())
    model = go.HeteroSGNNOracle(go.SAGEConv, torch.nn.ReLU(), 'sum', 128, 32, md, 2, 0.4, True,
                                False)
    with torch.no_grad():
        model(g.x_dict, ei)
    opt = torch.optim.Adam([p for p in model.parameters()
                            if not isinstance(p, torch.nn.parameter.UninitializedParameter)],
                           lr=0.01)
    y = g['artwork'].y_style
    times = []
    ref = None
    if parity:
        model.gnn.dropout_masks = parity_masks(g.num_nodes_dict)
        ref = {'state': {k: v.detach().clone() for k, v in model.state_dict().items()
                         if not isinstance(v, torch.nn.parameter.UninitializedParameter)}}
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        emb, out = model(g.x_dict, ei)
        loss = go.nll_loss_artwork(out[0], y)
        loss.backward()
        if parity and it == 0:
            ref.update(emb=emb['artwork'].detach().clone(), logp=out[0]['artwork'].detach().clone(),
                       loss=float(loss.item()),
                       grads={k: p.grad.detach().clone() for k, p in model.named_parameters()
                              if p.grad is not None})
            # second opinion for the gradients: the same step in float64.  Weight gradients are
            # sums over up to 116,475 rows; two float32 evaluations with different summation
            # orders (MKL sgemm here, split-K tensor-core tiles on the GPU) both sit 1e-5..1e-3
            # from the float64 value on the badly conditioned tensors, so "who is closer to
            # float64" is reported next to the raw difference
            import copy
            m64 = copy.deepcopy(model).double()
            m64.load_state_dict({k: v.double() for k, v in ref['state'].items()}, strict=False)
            m64.gnn.dropout_masks = {t: m.double() for t, m in model.gnn.dropout_masks.items()}
            m64.zero_grad(set_to_none=True)
            m64.train()
            _, out64 = m64({k: v.double() for k, v in g.x_dict.items()}, ei)
            go.nll_loss_artwork(out64[0], y).backward()
            ref['grads64'] = {k: p.grad.detach().clone() for k, p in m64.named_parameters()
                              if p.grad is not None}
            del m64, out64
        opt.step()
        loss.item()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return {'value': PASSES * n_edges * len(times) / total, 'ms_per_step': 1e3 * total / len(times),
            'n_edges': n_edges, 'n_artworks': int(g.num_nodes_dict['artwork']),
            'cores': torch.get_num_threads(), 'steps': len(times),
            'warmup': warmup, 'parity_ref': ref,
            'sample': f"synthetic ArtGraph '{size}' ({g.num_nodes_dict['artwork']} artworks, "
                      f"{n_edges} directed edges, one-hot features), {len(times)} train steps "
                      f"after {warmup} warm-up"}


def _rel(a, b):
    """max|a-b| / max|b| over the whole tensor (the metric of north_star's 1e-5, tests/util.py)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


def _grad_errors(ref_grads, named_params, ref64=None):
    """Per-tensor gradient error against ``ref_grads`` (name -> tensor): the worst max|a-b|/max|b|
    over the tensors above noise level (max|b| > 1e-4 of the model's largest gradient entry), its
    name, the worst max|a-b| over ALL tensors relative to the model's largest gradient, the
    number of tensors within 1e-5, and -- with the float64 gradients ``ref64`` -- the worst
    distance of the product and of the float32 reference from float64."""
    gmax = max(float(v.abs().max()) for v in ref_grads.values())
    worst, worst_name, glob, n, n_ok = 0.0, None, 0.0, 0, 0
    p64 = r64 = 0.0
    for k, gr in ref_grads.items():
        p = named_params.get(k)
        if p is None or p.grad is None:
            continue
        pg = p.grad.detach().double().cpu()
        err = float((pg - gr.double().cpu()).abs().max())
        glob = max(glob, err / gmax)
        scale = float(gr.abs().max())
        if scale <= 1e-4 * gmax:
            continue
        n += 1
        n_ok += err / scale <= 1e-5
        if err / scale > worst:
            worst, worst_name = err / scale, k
        if ref64 is not None and k in ref64:
            p64 = max(p64, float((pg - ref64[k]).abs().max()) / scale)
            r64 = max(r64, float((gr.double() - ref64[k]).abs().max()) / scale)
    out = {'grad_rel_err_max': worst, 'grad_worst_tensor': worst_name,
           'grad_err_over_model_grad_max': glob, 'grad_tensors': n, 'grad_tensors_within_1e-5': n_ok}
    if ref64 is not None:
        out['grad_product_vs_float64_max'] = p64
        out['grad_float32_reference_vs_float64_max'] = r64
    return out


def gpu_parity(ref, data, x, ei, y, dev):
    """One training step of the CUDA path from the oracle run's initial state with the same
    dropout masks (train_gnn_embeddings.py:39-52), compared with that run.  Outside every timed
    region; the checker is the oracle, the thing checked is the product."""
    import mmac_b200 as agx
    model = agx.HeteroSGNN(agx.SAGEConv, torch.nn.ReLU(), 'sum', 128, 32, data.metadata(), 2, 0.4,
                           True, False)
    missing, unexpected = model.load_state_dict(ref['state'], strict=False)
    assert not unexpected, unexpected
    model = model.to(dev).train()
    model.gnn.dropout_masks = {t: m.to(dev) for t, m in
                               parity_masks({t: v.shape[0] for t, v in x.items()}).items()}
    emb, out = model(x, ei)
    loss = agx.functional.nll_loss(out[0]['artwork'], y)
    loss.backward()
    torch.cuda.synchronize()
    gerr = _grad_errors(ref['grads'], dict(model.named_parameters()), ref.get('grads64'))
    return {'against': 'the cpu_baseline run of this very process (oracle port, same initial '
                       'weights, same injected dropout masks, first training step)',
            'metric': 'max|a-b| / max|b| per tensor', 'tolerance': 1e-5,
            'emb_rel_err': _rel(emb['artwork'], ref['emb']),
            'logp_rel_err': _rel(out[0]['artwork'], ref['logp']),
            'loss_rel_err': abs(float(loss.item()) - ref['loss']) / abs(ref['loss']), **gerr}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # bounded sample: one step is one full-size training step (~2.7 s on 16 cores, ~8 s on 8); at
    # most 12 timed steps after 1 warm-up keeps the arm within about half a minute -- the line
    # reports the steps and warm-up steps that actually ran
    r = cpu_reference(args.cpu_size, max(1, min(args.steps, 12)), 1)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': r['steps'], 'warmup': r['warmup'],
        'steps_requested': args.steps, 'warmup_requested': args.warmup,
        'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': _workload_name(args.size), 'operator': 'SAGEConv', 'label': 'style',
                   'hidden': 128, 'layers': 2, 'artworks_per_gpu': r['n_artworks'],
                   'directed_edges_per_gpu': r['n_edges'],
                   'edges_per_step_per_gpu': PASSES * r['n_edges'],
                   'parallelism': 'host CPU, all cores, one process',
                   'l2_policy': 'n/a (host CPU)', 'cuda_graph': False,
                   'note': 'PyG 2.0.2 semantics restated on torch ATen CPU kernels (oracle/); PyG '
                           'itself is not installable here'},
        'cpu_baseline': {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port',
                         'sample': r['sample']},
        'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def _reinit_for_parity(model, seed: int = 20261019):
    """Deterministic re-initialisation IN PLACE (the parameters are views of the optimizer's flat
    arena): U(+-1/sqrt(fan_in)) matrices, small biases, BatchNorm scale near 1, running statistics
    reset.  Host generator with a fixed seed: identical on every rank."""
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if isinstance(p, torch.nn.parameter.UninitializedParameter):
                continue
            u = torch.rand(p.shape, generator=gen) * 2 - 1
            if p.dim() >= 2:
                v = u / float(p.shape[1]) ** 0.5
            elif '.bns.' in name and name.endswith('weight'):
                v = 1.0 + 0.1 * u
            else:
                v = 0.1 * u
            p.copy_(v.to(p.device))
        for name, b in model.named_buffers():
            if name.endswith('running_mean'):
                b.zero_()
            elif name.endswith('running_var'):
                b.fill_(1.0)


def dist_parity_blocks(trainer, model, data, x, ei, y, world, rank, dev, dist, size='full'):
    """N > 1, block partition: one training step of the N-rank job (dropout masks injected,
    current weights) against the SAME step of the whole N-block graph on rank 0 alone -- what
    tests/test_gpu_dist.py checks on two GPUs, repeated here on every N the driver benchmarks
    (its own GPU test box has one GPU).  Errors are max|a-b| / max|b| per tensor."""
    import mmac_b200 as agx
    from mmac_b200 import synth
    from mmac_b200.dist import all_reduce_
    n_nodes = OrderedDict((t, int(v.shape[0])) for t, v in x.items())
    # Fresh random weights (same on every rank), not the state the timed steps left behind: the
    # check then does not depend on how long the timed loops trained (after a few hundred steps on
    # this synthetic task the loss is ~1e-5).  Gradient tensors are compared relative to their OWN
    # largest entry and, as a second figure, relative to the model's largest gradient: the small
    # tensors are sums of cancelling terms whose float32 value depends on the summation order
    # (N partial sums + NCCL here, one sum over N times the rows there) -- measured bit-identical
    # with and without side streams / in-place gradient accumulation (AGX_NO_SIDE_STREAMS,
    # AGX_NO_INPLACE_GRADS), i.e. arithmetic, not a race.
    _reinit_for_parity(model)
    state = {k: v.detach().clone() for k, v in model.state_dict().items()
             if not isinstance(v, torch.nn.parameter.UninitializedParameter)}
    model.gnn.dropout_masks = {t: m.to(dev) for t, m in
                               parity_masks(n_nodes, PARITY_MASK_SEED + rank).items()}
    model.train()
    trainer.opt.zero_grad()
    emb, out = model(trainer.x, ei)
    loss = agx.functional.nll_loss(out[0]['artwork'], trainer.y, dist.group.WORLD)
    loss.backward()
    all_reduce_(trainer.opt.grad, dist.group.WORLD)
    model.gnn.dropout_masks = None
    torch.cuda.synchronize()
    res = None
    if rank == 0:
        dense = agx.ToUndirected()(synth.make_artgraph(size, features='one-hot', seed=1234 + 2))
        whole = synth.replicate(dense.to(dev), world)
        ref = agx.HeteroSGNN(agx.SAGEConv, torch.nn.ReLU(), 'sum', 128, 32, data.metadata(), 2, 0.4,
                             True, False)
        ref.load_state_dict(state, strict=False)
        ref = ref.to(dev).train()
        ref.gnn.dropout_masks = {
            t: torch.cat([parity_masks(n_nodes, PARITY_MASK_SEED + q)[t] for q in range(world)]
                         ).to(dev) for t in n_nodes}
        emb_r, out_r = ref(whole.x_dict, whole.edge_index_dict)
        loss_r = agx.functional.nll_loss(out_r[0]['artwork'], whole['artwork'].y_style.to(dev))
        loss_r.backward()
        torch.cuda.synchronize()
        A = n_nodes['artwork']
        ref_grads = {k: p.grad for k, p in ref.named_parameters() if p.grad is not None}
        gerr = _grad_errors(ref_grads, dict(model.named_parameters()))
        res = {'against': f'the same training step of the whole {world}-block graph on ONE GPU '
                          f'(rank 0 alone, same weights, same injected dropout masks)',
               'metric': 'max|a-b| / max|b| per tensor', 'tolerance': 1e-5,
               'emb_rel_err': _rel(emb['artwork'], emb_r['artwork'][:A]),
               'logp_rel_err': _rel(out[0]['artwork'], out_r[0]['artwork'][:A]),
               'loss_rel_err': abs(float(loss.item()) - float(loss_r.item())) / abs(float(loss_r.item())),
               **gerr}
        del ref, whole, emb_r, out_r
    dist.barrier()
    return res


def config5_cut(world, rank, dev, dist, copies: int = 16, steps: int = 10, overlap: bool = True,
                scattered=()):
    """BASELINE configs[4]: ONE graph -- the 16x replicated synthetic ArtGraph with the artwork ids
    permuted, so that a contiguous cut crosses the copies -- partitioned by destination node over
    the N ranks (artwork rows cut; ``scattered`` types -- tag, artist -- cut into equal chunks, with a
    reduce-scatter of the artwork -> tag / artist partial sums to the owners and an all-gather of
    their rows for the reverse relations; every other node type replicated, the partial neighbour
    sums of the relations into them all-reduced inside every conv layer), against
    the same training step of the whole graph on one GPU (rank 0 alone).  Strong scaling:
    efficiency = t(1 GPU) / (N * t(N GPUs))."""
    import mmac_b200 as agx
    from mmac_b200 import synth
    from mmac_b200.dist import GraphPartition, partition_context
    from mmac_b200.trainer import GNNTrainer
    g = synth.make_artgraph('full', features='dense', seed=1234 + 5)
    data = agx.ToUndirected()(g)
    whole = synth.replicate(data.to(dev), copies)
    n = whole.num_nodes_dict
    A = n['artwork']
    perm = torch.randperm(A, generator=torch.Generator().manual_seed(55)).to(dev)   # old -> new id
    x = OrderedDict(whole.x_dict)
    xa = torch.empty_like(x['artwork'])
    xa[perm] = x['artwork']
    x['artwork'] = xa
    y = torch.empty(A, dtype=torch.int64, device=dev)
    y[perm] = whole['artwork'].y_style.to(torch.int64)
    ei = OrderedDict()
    for (s, r, d), e in whole.edge_index_dict.items():
        ei[(s, r, d)] = torch.stack([perm[e[0]] if s == 'artwork' else e[0],
                                     perm[e[1]] if d == 'artwork' else e[1]]).contiguous()
    total_edges = sum(int(v.shape[1]) for v in ei.values())
    scattered = [t for t in scattered if t in n and n[t] >= world]
    part = GraphPartition(ei, n, world, rank, scattered=scattered,
                          replicated=[t for t in n if t != 'artwork' and t not in scattered])
    ctx = partition_context(part, dist.group.WORLD, dev)
    torch.manual_seed(0)
    model = agx.HeteroSGNN(agx.SAGEConv, torch.nn.ReLU(), 'sum', 128, 32, data.metadata(), 2, 0.4,
                           True, False).to(dev)
    x_own = OrderedDict((t, part.owned(t, v).contiguous()) for t, v in x.items())
    tr = GNNTrainer(model, x_own, part.edge_index, part.owned('artwork', y), lr=0.01,
                    use_cuda_graph=True, dist_ctx=ctx)

    def timed(trainer, k):
        for _ in range(3):
            trainer.train_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            loss_ = trainer.train_step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k, float(loss_.item())

    dist.barrier()
    ms_n, loss_n = timed(tr, steps)
    t = torch.tensor([ms_n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_n = float(t.item())
    # bytes every rank contributes to the per-layer all-reduce of partial neighbour sums (forward;
    # the backward all-reduce of their gradients moves the same amount for the layers whose input
    # needs a gradient)
    part_bytes = []
    for width in (128, 128, 128):            # conv0 (dense 128-d inputs), conv1, conv_out inputs
        part_bytes.append(sum(n[d] * width * 4 for (s, r, d) in part.partial))
    scat_bytes = sum(int(c.shape[0]) * 128 * 4 for c in part.scatter.values())
    res = None
    del tr
    if rank == 0:
        torch.manual_seed(0)
        m1 = agx.HeteroSGNN(agx.SAGEConv, torch.nn.ReLU(), 'sum', 128, 32, data.metadata(), 2, 0.4,
                            True, False).to(dev)
        t1 = GNNTrainer(m1, x, ei, y, lr=0.01, use_cuda_graph=True)
        ms_1, loss_1 = timed(t1, steps)
        res = {'workload': f"{copies}x replicated synthetic ArtGraph 'full' (128-d features for every "
                           f"node type), artwork ids permuted: {A} artworks, {total_edges} directed "
                           f"edges; SAGEConv training step",
               'partition': f'artwork rows cut into {world} contiguous ranges' +
                            (f'; {", ".join(scattered)} cut into {world} equal chunks (partial sums '
                             f'into them reduce-scattered to the owners, their rows all-gathered '
                             f'for the reverse relations)' if scattered else '') +
                            '; the other node types replicated, per conv layer ONE all-reduce of '
                            'the partial neighbour sums of the relations into them',
               'scattered_types': list(scattered),
               'reduce_scatter_bytes_per_layer_fwd': scat_bytes,
               'n_gpus': world, 'ms_per_step': ms_n, 'ms_per_step_1gpu': ms_1,
               'edges_per_s': PASSES * total_edges / (ms_n * 1e-3),
               'edges_per_s_1gpu': PASSES * total_edges / (ms_1 * 1e-3),
               'speedup_vs_1gpu': ms_1 / ms_n, 'strong_scaling_efficiency': ms_1 / (world * ms_n),
               'allreduce_bytes_per_layer_fwd': part_bytes,
               'boundary_rows_all_gathered': part.halo_rows(),
               'edges_on_rank0': sum(int(v.shape[1]) for v in part.edge_index.values()),
               'final_loss_Ngpu': loss_n, 'final_loss_1gpu': loss_1}
        del t1, m1
    dist.barrier()
    return res


def heads_algorithmic_bytes(batch, fv=768, emb=128, classes=(32, 18)):
    """HBM bytes of one multitask head training step (forward AND backward): the
    inputs (feat + two embeddings) are read once by the forward and once more by the weight-gradient
    pass, logits written and their gradient read, weights read and their gradient written."""
    c = sum(classes)
    return 2 * batch * (fv + 2 * emb) * 4 + 2 * batch * c * 4 + 2 * c * (fv + emb) * 4


def heads_cpu_reference(batch, steps=20):
    """The reference's head arithmetic (oracle/heads_oracle.py: cat -> Dropout -> Linear, weighted CE,
    Adam) on the host cores, float32 -- the cpu_baseline of the heads half."""
    from mmac_b200 import synth
    from oracle import heads_oracle as ho
    torch.set_num_threads(os.cpu_count() or 1)
    feat, es, eg, ys, yg = synth.make_head_batch(batch, 'vit', seed=1)
    ws, wg = synth.class_weights(ys, 32), synth.class_weights(yg, 18)
    m = ho.MultiTaskHeadOracle(768, 128, {'style': 32, 'genre': 18}, 0.4).train()
    opt = torch.optim.Adam(m.parameters(), lr=3e-4)
    ts = []
    for it in range(steps + 2):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = ho.multitask_loss(m(feat, es, eg), ys, yg, ws, wg)
        loss.backward()
        opt.step()
        loss.item()
        if it >= 2:
            ts.append(time.perf_counter() - t0)
    return {'value': batch * len(ts) / sum(ts), 'unit': 'artworks/s', 'cores': torch.get_num_threads(),
            'kind': 'port', 'sample': f'multitask ViT heads, batch {batch}, {len(ts)} train steps '
                                      f'after 2 warm-up, float32'}


def heads_throughput(dev, dist, world, batches=(4096,), steps: int = 20, cpu: bool = False):
    """Secondary metric of north_star (artworks/s): one training step (forward + loss + backward +
    Adam) of the new-multimodal multitask ViT heads (configs[3], reference batch loop
    src/train_new_multimodal_multitask.py:62-90) and of the projector (configs[2],
    src/train_projector.py:39-59) on precomputed synthetic features, batch-sharded over the ranks
    with one gradient all-reduce per step; in bf16 on the tensor cores (the fused agx_head_step: the
    precision class of the reference's fp16 autocast) and in exact float32.  Device-resident
    (a ring of distinct batches larger than the L2, so inputs come from HBM) and end to end
    (pinned host -> device copy of the batch, loss read-back)."""
    import mmac_b200 as agx
    from mmac_b200 import synth
    from mmac_b200.trainer import HeadTrainer
    group = dist.group.WORLD if dist is not None else None
    rank = int(os.environ.get('RANK', '0'))
    peak, _ = _peaks()
    out = {'features': 'synthetic ViT CLS features [B,768] + style/genre embeddings [B,128]',
           'l2_policy': 'device-resident leg cycles through a ring of distinct batches of >= 160 MB '
                        'in total (> 126 MB L2)', 'runs': []}
    for batch in batches:
        in_bytes = batch * (768 + 256) * 4
        n_ring = int(min(64, max(2, -(-160_000_000 // in_bytes))))
        ring = []
        for i in range(n_ring):
            ring.append([t.to(dev) for t in synth.make_head_batch(batch, 'vit', seed=1 + 97 * rank + i)])
        host = [t.pin_memory() for t in synth.make_head_batch(batch, 'vit', seed=1 + 97 * rank)]
        ws = synth.class_weights(host[3], 32).to(dev)
        wg = synth.class_weights(host[4], 18).to(dev)
        for precision in ('bf16', 'fp32'):
            for kind in ('multitask', 'projector'):
                torch.manual_seed(0)
                if kind == 'multitask':
                    head = agx.NewMultiModalMultiTaskHead(128, {'style': 32, 'genre': 18}, 0.4, 768).to(dev)
                    tr = HeadTrainer(head, 'multitask', 3e-4, ws, wg, group=group, use_cuda_graph=True,
                                     precision=precision)
                    sel = lambda b: b                                   # noqa: E731
                else:
                    head = agx.LabelProjectorHead(128, 768).to(dev)
                    tr = HeadTrainer(head, 'projector', 3e-4, group=group, use_cuda_graph=True,
                                     precision=precision)
                    sel = lambda b: [b[0], b[1]]                        # noqa: E731
                for i in range(3):
                    tr.step(*sel(ring[i % n_ring]))
                torch.cuda.synchronize()
                if dist is not None:
                    dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(steps):
                    tr.step(*sel(ring[i % n_ring]))
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                # end to end: every step's batch travels from pinned host memory (copy stream, one
                # step ahead: HeadTrainer.prefetch / step_prefetched), the loss of EVERY step is
                # read on the host one step late
                pin = [torch.empty(1, dtype=torch.float32, pin_memory=True) for _ in range(2)]
                evs = [torch.cuda.Event() for _ in range(2)]

                def e2e_loop(n):
                    tr.prefetch(*sel(host))
                    last = None
                    for i in range(n):
                        loss = tr.step_prefetched()
                        if i + 1 < n:
                            tr.prefetch(*sel(host))
                        pin[i & 1].copy_(loss.detach().reshape(1), non_blocking=True)
                        evs[i & 1].record()
                        if i > 0:
                            evs[(i - 1) & 1].synchronize()
                            last = float(pin[(i - 1) & 1][0])
                    evs[(n - 1) & 1].synchronize()
                    return float(pin[(n - 1) & 1][0])
                e2e_loop(2)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                e2e_loop(steps)
                torch.cuda.synchronize()
                ms_e2e = (time.perf_counter() - t0) * 1e3
                if dist is not None:
                    t = torch.tensor([ms, ms_e2e], device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms, ms_e2e = float(t[0]), float(t[1])
                run = {'kind': kind, 'precision': precision, 'batch_per_gpu': batch,
                       'ms_per_step': ms / steps,
                       'artworks_per_s': world * batch * steps / (ms * 1e-3),
                       'e2e_artworks_per_s': world * batch * steps / (ms_e2e * 1e-3)}
                if kind == 'multitask':
                    nb = heads_algorithmic_bytes(batch)
                    gbs = nb / (ms / steps * 1e-3) / 1e9
                    run['roofline'] = {'bound': 'hbm', 'achieved': gbs, 'peak': peak, 'unit': 'GB/s',
                                       'frac': gbs / peak, 'bytes_per_step': nb,
                                       'what': 'whole step (copy of the batch into the static '
                                               'buffers, prepare + fused + reduce kernels, Adam) '
                                               'under CUDA-graph replay'}
                out['runs'].append(run)
                del tr, head
        del ring
    # the summary the driver's earlier lines carried (batch 4096, tensor-core path)
    for r in out['runs']:
        if r['batch_per_gpu'] == 4096 and r['precision'] == 'bf16':
            out[r['kind']] = {k: r[k] for k in ('artworks_per_s', 'e2e_artworks_per_s', 'ms_per_step')}
            if 'roofline' in r:
                out['roofline'] = r['roofline']
    out['batch_per_gpu'] = 4096
    if cpu and rank == 0:
        out['cpu_baseline'] = heads_cpu_reference(4096)
    return out


def operator_throughput(opname, metadata, x, ei, y, dev, n_edges, steps: int = 10):
    """Secondary line: the same training step with another operator of the reference's
    ``operator_registry`` (train_gnn_embeddings.py:96-102; GATConv is the script's default
    ``--operator``), inputs resident, CUDA-graph replay, timed with CUDA events."""
    import mmac_b200 as agx
    from mmac_b200.trainer import GNNTrainer
    torch.manual_seed(0)
    model = agx.HeteroSGNN(getattr(agx, opname), torch.nn.ReLU(), 'sum', 128, 32, metadata, 2, 0.4,
                           True, False).to(dev)
    tr = GNNTrainer(model, x, ei, y, lr=0.01, use_cuda_graph=True)
    for _ in range(3):
        tr.train_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = tr.train_step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {'ms_per_step': ms, 'edges_per_s': PASSES * n_edges / (ms * 1e-3),
            'launches_per_step': int(tr.launches_per_step), 'loss': float(loss.item())}


def _workload_name(size, operator='SAGEConv'):
    return (f"train_gnn_embeddings.py --label style: full-graph {operator} to_hetero training step on "
            f"synthetic ArtGraph '{size}' (one-hot node features as in artgraph.py:93-95)")


# ------------------------------------------------------------------------------------------------
# this repo's CUDA path
def _pin_to_gpu_cpus(local: int):
    """One process per GPU on a two-socket host: run this rank (and allocate its pinned staging
    buffers, first-touch) on the cores NVML lists as local to its GPU -- what ``numactl
    --cpunodebind`` would do in a launcher script.  Returns the number of cores, or None when NVML
    gives nothing usable (nothing is changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if len(cpus) < 2:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import mmac_b200 as agx
    from mmac_b200 import ops, synth
    from mmac_b200.trainer import GNNTrainer
    from mmac_b200._lib import launch_count

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product has no CPU path '
                         '(use --impl reference for the host baseline)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa = _pin_to_gpu_cpus(local) if (world > 1 and not args.no_numa_pin) else None
    dist = None
    if world > 1:
        # the image exports NCCL_DEBUG=VERSION, which makes NCCL print a banner on stdout next to
        # the one JSON line this script owes the driver: whatever native libraries write to fd 1
        # goes to stderr (NCCL's banner stays visible there); the JSON line is written to the
        # saved original stdout
        global _JSON_OUT
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    agx.lib()
    # ---- inputs: pinned host copies (e2e) + device residents ---------------------------------
    cut = args.partition == 'cut' and world > 1
    part = None
    if cut:
        # ONE graph cut into `world` destination ranges per node type (strong scaling): edges cross
        # ranks, so every conv layer all-gathers boundary source rows and the backward pass
        # reduce-scatters their gradients.  128-d features for every node type (SURVEY.md 8d
        # variant): boundary rows of one-hot inputs would be N_type floats wide.
        from mmac_b200.dist import GraphPartition, balanced_bounds
        g = synth.make_artgraph(args.size, features='dense')
        data = agx.ToUndirected()(g)
        # equal-rows ranges by default; --balanced-cut equalises incoming edges per rank instead
        # (measured at 2 GPUs: 4.75 ms/step equal rows, 7.37 ms/step equal edges -- the hub
        # relations dominate either way)
        bounds = balanced_bounds(data.edge_index_dict, data.num_nodes_dict, world) \
            if args.balanced_cut else None
        # --replicate-small: every node type but artwork lives on all ranks: no
        # boundary rows, the artwork -> X neighbour sums are all-reduced inside the conv layer
        part = GraphPartition(data.edge_index_dict, data.num_nodes_dict, world, rank, bounds,
                              replicated=[t for t in data.num_nodes_dict if t != 'artwork']
                              if args.replicate_small else ())
        host_x = OrderedDict((k, part.owned(k, v).contiguous().pin_memory())
                             for k, v in data.x_dict.items())
        host_ei = OrderedDict((k, v.pin_memory()) for k, v in part.edge_index.items())
        y_all = part.owned('artwork', data['artwork'].y_style)
        total_edges = sum(int(v.shape[1]) for v in data.edge_index_dict.values())
    else:
        # the reference's one-hot features torch.eye(N_t) (artgraph.py:93-95) are passed as
        # agx.Identity(N_t) markers: the same inputs without building or uploading the N x N
        # matrices (148 of 236 MB per step); --dense-onehot passes the dense matrices, which the
        # module then detects on the device (the round-1 behaviour)
        g = synth.make_artgraph(args.size, features='one-hot' if args.dense_onehot else 'identity',
                                seed=None if world == 1 else 1234 + 2)
        data = agx.ToUndirected()(g)
        host_x = OrderedDict((k, v.pin_memory() if torch.is_tensor(v) else v)
                             for k, v in data.x_dict.items())
        host_ei = OrderedDict((k, v.pin_memory()) for k, v in data.edge_index_dict.items())
        y_all = data['artwork'].y_style
        total_edges = world * sum(int(v.shape[1]) for v in host_ei.values())
    x = OrderedDict((k, v.to(dev, non_blocking=True) if torch.is_tensor(v) else v)
                    for k, v in host_x.items())
    ei = OrderedDict((k, v.to(dev, non_blocking=True)) for k, v in host_ei.items())
    y = y_all.to(dev)
    n_edges = sum(int(v.shape[1]) for v in ei.values())
    h2d = sum(v.numel() * v.element_size() for v in host_x.values() if torch.is_tensor(v)) + \
        sum(v.numel() * v.element_size() for v in host_ei.values())

    torch.manual_seed(0)
    model = agx.HeteroSGNN(getattr(agx, args.operator), torch.nn.ReLU(), 'sum', 128, 32,
                           data.metadata(), 2, 0.4, True, False).to(dev)
    ctx = None
    if dist is not None:
        # the world-times replicated block-diagonal graph (BASELINE config 5), one block per rank:
        # no edge is cut, so no feature rows move; BatchNorm statistics, the loss and the weight
        # gradients are those of the whole graph (NCCL all-reduces inside the captured step)
        from mmac_b200.dist import block_context, partition_context
        ctx = partition_context(part, dist.group.WORLD, dev) if cut else \
            block_context(dist.group.WORLD, {t: int(v.shape[0]) for t, v in x.items()})
    trainer = GNNTrainer(model, x, ei, y, lr=0.01, use_cuda_graph=not args.no_graph, dist_ctx=ctx)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if args.ncu_heads:
        # profiling aid: exactly one eager fused head step (B = 4096, bf16) between
        # cudaProfilerStart / Stop
        from mmac_b200.trainer import HeadTrainer
        hb = [t.to(dev) for t in synth.make_head_batch(4096, 'vit', seed=1)]
        hws, hwg = synth.class_weights(hb[3].cpu(), 32).to(dev), synth.class_weights(hb[4].cpu(), 18).to(dev)
        hh = agx.NewMultiModalMultiTaskHead(128, {'style': 32, 'genre': 18}, 0.4, 768).to(dev)
        htr = HeadTrainer(hh, 'multitask', 3e-4, hws, hwg, use_cuda_graph=False, precision='bf16')
        ph = agx.LabelProjectorHead(128, 768).to(dev)
        ptr_ = HeadTrainer(ph, 'projector', 3e-4, use_cuda_graph=False, precision='bf16')
        for _ in range(3):
            htr.step(*hb)
            ptr_.step(hb[0], hb[1])
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        htr.step(*hb)
        ptr_.step(hb[0], hb[1])
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        print(json.dumps({'ncu_heads_done': True}))
        return
    if args.ncu:
        # profiling aid: `ncu --profile-from-start off ... bench.py --ncu` sees exactly one eager
        # training step (every kernel launched individually, no CUDA graph)
        trainer.use_cuda_graph = False
        for _ in range(3):
            trainer.train_step()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        trainer.train_step()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        print(json.dumps({'ncu_step_done': True, 'loss': float(trainer._loss.item())}))
        return

    # ---- device-resident timing --------------------------------------------------------------
    # the clock sampler (nvidia-smi, one sample per 20 ms) needs ~0.1 s to come up and the timed
    # region is 20 x 2 ms: it is started before the warm-up steps and the GPU is kept under the same
    # load (untimed steps) until it has delivered samples; it stops right after the timed region
    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(max(args.warmup, 3)):
        trainer.train_step()
    for _ in range(150):                # untimed steps under the sampler (a FIXED count: every
        trainer.train_step()            # step holds collectives, all ranks must run the same steps)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = launch_count()
    ev0.record()
    for _ in range(args.steps):
        loss = trainer.train_step()
    ev1.record()
    barrier()
    clk = clocks.stop()
    ms = ev0.elapsed_time(ev1)
    launches = (launch_count() - l0) if args.no_graph else trainer.launches_per_step * args.steps
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = PASSES * total_edges * args.steps / (ms * 1e-3)
    final_loss = float(loss.item())

    # ---- end to end: host buffers in, loss out, every step ------------------------------------
    # Every step: its inputs travel from pinned host memory to the device (prefetched one step
    # ahead on a copy stream, so the PCIe transfer overlaps the previous step's kernels), are moved
    # into the captured step's static tensors, the CSR/CSC are re-sorted, the step runs, the loss
    # is read back.  `--e2e-serial` copies in line instead (no overlap).
    # The loss of EVERY step is read back (4 bytes D2H into pinned memory) and the input checks of
    # every step are evaluated on the host -- one step late: step i+1 is queued before the host
    # waits for step i, so the GPU never idles while Python enqueues (--e2e-serial: wait in line).
    loss_pin = [torch.empty(1, dtype=torch.float32, pin_memory=True) for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]

    def e2e_enqueue(i, first):
        if args.e2e_serial:
            trainer.update_inputs(host_x, host_ei)      # H2D of x_dict + edge_index_dict, re-sort
        else:
            if first:
                trainer.prefetch_inputs(host_x, host_ei)
            trainer.consume_prefetched()                # D2D into the static tensors, re-sort
            trainer.prefetch_inputs(host_x, host_ei)    # next step's H2D, overlaps this step
        loss_dev = trainer.train_step()
        loss_pin[i & 1].copy_(loss_dev.reshape(1), non_blocking=True)      # D2H of the loss
        loss_ev[i & 1].record()
        return trainer.verify_inputs_async()

    def e2e_finish(i, check):
        loss_ev[i & 1].synchronize()
        check()
        return float(loss_pin[i & 1][0])

    def e2e_run(n, first):
        prev = None
        last = None
        for i in range(n):
            chk = e2e_enqueue(i, first and i == 0)
            if args.e2e_serial:
                last = e2e_finish(i, chk)
            else:
                if prev is not None:
                    last = e2e_finish(*prev)
                prev = (i, chk)
        if prev is not None:
            last = e2e_finish(*prev)
        return last

    e2e_run(2, True)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    l_host = e2e_run(args.steps, False)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(e2e_ms, wall_ms)
    if dist is not None:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = PASSES * total_edges * args.steps / (e2e_ms * 1e-3)

    # ---- roofline of the aggregation kernels: events around every launch, eager steps ----------
    roofline = None
    cpu_base = None
    parity = None
    timers = [ops.KernelTimer() for _ in range(3)]
    for timer in timers:                 # every rank steps (the step holds collectives)
        if rank == 0:
            ops.TIMER = timer
        # an eager step is CPU-bound (~150 launches from Python): park the GPU behind a spin kernel
        # while the CPU enqueues the whole step, so that the CUDA events around each launch measure
        # kernel time, not the gaps in which the GPU waits for the next launch
        torch.cuda.synchronize()
        torch.cuda._sleep(int(args.park_ms * 1e-3 * 1.9e9))
        trainer._step_eager()
    ops.TIMER = None
    barrier()
    if rank == 0:
        peak, peak_src = _peaks()
        # every launch position is timed in three steps: keep its fastest time (a step whose
        # enqueue outlasted the parking delay shows launch gaps, not kernel time)
        summ = ops.KernelTimer.summary_min(timers)
        agg = {k: v for k, v in summ.items() if k.startswith('agg')}
        if not agg:      # GATConv: its attention kernels are not instrumented (not the headline path)
            agg = {'none': {'launches': 1, 'ms': float('inf'), 'bytes': 0, 'flops': 0}}
        dom = max(agg, key=lambda k: agg[k]['ms'])
        d = agg[dom]
        achieved = d['bytes'] / (d['ms'] * 1e-3) / 1e9
        all_b = sum(v['bytes'] for v in agg.values())
        all_ms = sum(v['ms'] for v in agg.values())
        # three readings of the same number: against the measured copy peak, against
        # north_star's nominal "about 8 TB/s", and against what the L2 delivers for this access
        # pattern (profiles/probes/gather_probe.cu: 15.6 TB/s of random 512 B rows out of an
        # L2-resident 60 MB table) -- the source tables of these launches live in L2, so the last
        # one is the physical ceiling
        roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak,
                    'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None,
                    'peak_source': peak_src, 'launches_timed': d['launches'],
                    'avg_launch_us': 1e3 * d['ms'] / d['launches'],
                    'timing': 'CUDA events around every launch of three eager steps queued behind '
                              'a parked GPU; per launch position the fastest of the three (min of 3)',
                    'frac_of_nominal_8000': achieved / 8000.0,
                    'frac_of_l2_gather_ceiling_15600': achieved / 15600.0,
                    'per_kernel': {k: {'launches': v['launches'], 'us_per_step': 1e3 * v['ms'],
                                       'achieved': v['bytes'] / (v['ms'] * 1e-3) / 1e9,
                                       'frac': v['bytes'] / (v['ms'] * 1e-3) / 1e9 / peak,
                                       'frac_of_nominal_8000': v['bytes'] / (v['ms'] * 1e-3) / 8e12}
                                   for k, v in agg.items() if v['bytes']},
                    'all_aggregation': {'achieved': all_b / (all_ms * 1e-3) / 1e9,
                                        'frac': all_b / (all_ms * 1e-3) / 1e9 / peak,
                                        'frac_of_nominal_8000': all_b / (all_ms * 1e-3) / 8e12,
                                        'ms_per_step': all_ms,
                                        'bytes_per_step': all_b},
                    'gemm': ({'tflops': summ['gemm']['flops'] / (summ['gemm']['ms'] * 1e-3) / 1e12,
                              'ms_per_step': summ['gemm']['ms']} if 'gemm' in summ else None)}
        tr = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tr):
            with open(tr) as fh:
                tj = json.load(fh)
            roofline['traffic'] = tj.get(dom)
            # NOT measured by this run: ncu cannot run inside the timed process
            roofline['traffic_source'] = tj.get('_source', 'profiles/traffic.json (ncu --set full '
                                                           'capture, see profiles/)')
        if world == 1 and not args.no_cpu_baseline:
            want_parity = args.cpu_size == args.size and args.operator == 'SAGEConv'
            r = cpu_reference(args.cpu_size, 2, 1, parity=want_parity)
            cpu_base = {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port',
                        'sample': r['sample']}
            if want_parity:
                parity = gpu_parity(r['parity_ref'], data, x, ei, y, dev)

    # the reference's epoch (train_gnn_embeddings.py:149-151): one training step + hetero_test(),
    # i.e. two evaluation forwards with loss and accuracy (validation and test graph; here the
    # synthetic training graph stands in for both -- same node and edge counts)
    epoch = None
    try:
        for _ in range(2):
            trainer.train_step(); trainer.evaluate(); trainer.evaluate()
        torch.cuda.synchronize()
        barrier()
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ee0.record()
        n_ep = 10
        for _ in range(n_ep):
            trainer.train_step()
            trainer.evaluate()
            ev_loss, ev_acc = trainer.evaluate()
        ee1.record()
        torch.cuda.synchronize()
        ep_ms = ee0.elapsed_time(ee1) / n_ep
        if dist is not None:
            t = torch.tensor([ep_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ep_ms = float(t.item())
        epoch = {'ms': ep_ms, 'what': 'training step + 2 evaluation forwards (loss + accuracy), all '
                                      'three CUDA-graph replays, inputs resident',
                 'eval_loss': float(ev_loss.item()), 'eval_accuracy': float(ev_acc.item())}
    except Exception as e:  # noqa: BLE001  (a secondary figure must not cost the headline)
        epoch = {'error': f'{type(e).__name__}: {e}'[:300]}
    dist_parity = None
    config5 = None
    if dist is not None and not cut and args.operator == 'SAGEConv':
        if not args.no_dist_parity:
            dist_parity = dist_parity_blocks(trainer, model, data, x, ei, y, world, rank, dev, dist, args.size)
        if not args.no_config5:
            del trainer
            torch.cuda.empty_cache()
            config5 = config5_cut(world, rank, dev, dist, copies=args.config5_copies,
                                  scattered=[t for t in args.config5_scattered.split(',') if t])
    heads = None
    if not args.no_heads:
        heads = heads_throughput(dev, dist, world, batches=(32, 4096, 65536) if world == 1 else (4096,),
                                 cpu=(world == 1 and not args.no_cpu_baseline))
    operators = None
    if world == 1 and not args.no_operators and args.operator == 'SAGEConv':
        operators = {}
        for opname in ('GraphConv', 'GATConv'):
            try:
                operators[opname] = operator_throughput(opname, data.metadata(), x, ei, y, dev,
                                                        n_edges)
            except Exception as e:  # noqa: BLE001  (a secondary line must not cost the headline)
                operators[opname] = {'error': f'{type(e).__name__}: {e}'[:300]}
    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'strong' if cut else 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': _workload_name(args.size, args.operator) if not cut else
                       _workload_name(args.size, args.operator).replace('one-hot node features as in '
                                                         'artgraph.py:93-95',
                                                         '128-d features for every node type'),
                       'operator': args.operator,
                       'label': 'style', 'hidden': 128, 'layers': 2,
                       'artworks_per_gpu': int(x['artwork'].shape[0]),
                       'one_hot_features': 'dense torch.eye matrices' if (cut or args.dense_onehot) else
                                           'agx.Identity(n) markers (no N x N matrix built or uploaded)',
                       'directed_edges_per_gpu': n_edges,
                       'edges_per_step_per_gpu': PASSES * n_edges,
                       'parallelism': 'single GPU' if world == 1 else (
                           f'artwork rows cut into {world} ranges, every other node type replicated '
                           f'on all ranks; per conv layer one all-reduce of the partial neighbour '
                           f'sums of the artwork -> X relations (and one of their gradients); '
                           f'weight-gradient, BatchNorm and loss all-reduces'
                           if cut and args.replicate_small else
                           f'one graph cut into {world} destination ranges per node type; per conv '
                           f'layer an NCCL all-gather of boundary source rows '
                           f'({part.halo_rows()} rows received per exchange on rank 0) and a '
                           f'reduce-scatter of their gradients; weight-gradient, BatchNorm and loss '
                           f'all-reduces' if cut else
                           f'{world} graph blocks, one per GPU; weight-gradient + '
                           f'BatchNorm-statistic all-reduce (NCCL)'),
                       'l2_policy': 'no explicit flush: one step streams ~1 GB of activations, '
                                    'gradients and workspaces (8x the 126 MB L2) between two uses of '
                                    'any buffer; the 60 MB artwork table and the small tables DO stay '
                                    'L2-resident inside an aggregation launch (roofline.traffic)',
                       'cuda_graph': not args.no_graph, 'final_loss': final_loss},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': 4, 'ms_per_step': e2e_ms / args.steps,
                    'includes': 'per step: pinned-host -> device copy of x_dict and edge_index_dict '
                                + ('(in line)' if args.e2e_serial else
                                   '(prefetched one step ahead on a copy stream, overlapping the '
                                   'previous step)') +
                                ', CSR/CSC re-sort, train step, loss read-back and input checks '
                                'of every step (evaluated on the host one step late, so step i+1 '
                                'is queued while step i runs)',
                    'last_loss': l_host},
            'gpu_launches': int(launches),
            'clocks': clk,
            'roofline': roofline,
            'cpu_baseline': cpu_base,
            'parity': parity,
            'dist_parity': dist_parity,
            'small_allreduce': (__import__('mmac_b200.dist', fromlist=['PEER_STATUS']).PEER_STATUS
                                if dist is not None else None),
            'host_cores_bound': numa,       # N > 1: cores local to the rank's GPU (NVML), or None
            'config5': config5,
            'epoch_ms': epoch.get('ms') if epoch else None,
            'epoch': epoch,
            'heads': heads,
            'operators': operators,
        }
        print(json.dumps(line), file=_JSON_OUT or sys.stdout)
        (_JSON_OUT or sys.stdout).flush()
    if dist is not None:
        # captured graphs hold NCCL work: finish everything, then leave without tearing the
        # communicator down under them
        barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--size', default='full', help="synthetic graph size of the GPU arm")
    ap.add_argument('--cpu-size', default='full',
                    help="graph size of the bounded CPU sample (full: ~8 s per step on 8 cores)")
    ap.add_argument('--operator', default='SAGEConv', choices=['SAGEConv', 'GraphConv', 'GATConv'],
                    help='conv operator of the GPU arm (the headline configuration is SAGEConv)')
    ap.add_argument('--partition', default='blocks', choices=['blocks', 'cut'],
                    help="N > 1: 'blocks' = N-times replicated graph, one block per rank (weak "
                         "scaling, the default the driver measures); 'cut' = one graph cut by "
                         "destination node with boundary-row all-gather (strong scaling)")
    ap.add_argument('--replicate-small', action='store_true',
                    help="--partition cut: replicate every node type but artwork on all ranks "
                         "(partial neighbour sums all-reduced; no boundary-row exchange)")
    ap.add_argument('--balanced-cut', action='store_true',
                    help="--partition cut: destination ranges with equal incoming edges, not rows")
    ap.add_argument('--dense-onehot', action='store_true',
                    help='pass the one-hot features as dense torch.eye matrices (uploaded every '
                         'e2e step: 236 MB) instead of agx.Identity markers (88 MB)')
    ap.add_argument('--e2e-serial', action='store_true',
                    help='end-to-end leg: copy the inputs in line instead of prefetching them')
    ap.add_argument('--no-graph', action='store_true', help='eager launches instead of CUDA graph')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-heads', action='store_true',
                    help='skip the secondary fusion-head / projector artworks/s measurement')
    ap.add_argument('--no-dist-parity', action='store_true',
                    help='N > 1: skip the N-rank step vs whole-graph-on-one-GPU comparison')
    ap.add_argument('--no-config5', action='store_true',
                    help='N > 1: skip the strong-scaling measurement of the cut 16x graph')
    ap.add_argument('--no-numa-pin', action='store_true',
                    help='N > 1: do not bind the rank to the cores local to its GPU')
    ap.add_argument('--config5-copies', type=int, default=16)
    ap.add_argument('--config5-scattered', default='',
                    help="node types cut into equal chunks in the config-5 partition, e.g. "
                         "'tag,artist' (default '': replicate everything but artwork -- measured "
                         "faster at 2 and at 8 GPUs: 13.3 vs 15.0 ms, 5.67 vs 5.85 ms)")
    ap.add_argument('--no-operators', action='store_true',
                    help='skip the secondary GraphConv / GATConv training-step measurement')
    ap.add_argument('--park-ms', type=float, default=120.0,
                    help='device-side delay in front of each per-kernel timing step (roofline leg)')
    ap.add_argument('--ncu-heads', action='store_true',
                    help='run one eager bf16 head step between cudaProfilerStart/Stop and exit')
    ap.add_argument('--ncu', action='store_true',
                    help='run one eager step between cudaProfilerStart/Stop and exit')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()

"""CPU suite, part 4: HOST-SIDE logic of the fused layers and of the whole hetero model (functional.
_HeteroGATFn / _HeteroConvFn / _BNActFn / losses, hetero.HeteroModule): which buffers meet which index
arrays, the gradient formulas, the GEMM wave order, descriptor wiring -- checked against the oracle with
every agx entry point restated in torch (tests/cpu_shim.py).  The kernels themselves are compared with
the oracle on the GPU (tests/test_gpu_model.py, test_gpu_kernels.py)."""
from collections import OrderedDict

import pytest
import torch

import util
import mmac_b200 as agx
from cpu_shim import cpu_ops
from oracle import graph_oracle as go
from util import rel_err


def _pair(C, xs, xd, ei, same):
    orc = go.GATConv((-1, -1), C)
    with torch.no_grad():
        orc(xs if same else (xs, xd), ei)
    util.fill_params_deterministic(orc)
    with torch.no_grad():
        orc.bias.copy_(torch.linspace(-0.3, 0.4, C))
    prod = agx.GATConv((-1, -1), C)
    util.copy_state(orc, prod)
    return orc.double(), prod


CASES = [  # n_src, n_dst, edges, F_src, F_dst, C, same tensor, hub
    (70, 50, 400, 24, 40, 32, False, False),
    (60, 60, 300, 16, 16, 128, True, False),
    (9, 300, 500, 8, 12, 18, False, False),
    (400, 6, 5000, 12, 10, 32, False, True),       # rows of ~800 edges, one of > 1024
]


@pytest.mark.parametrize('case', CASES)
def test_gatconv_host_logic_vs_oracle(case):
    n_src, n_dst, e, fs, fd, C, same, hub = case
    gen = torch.Generator().manual_seed(21)
    src = torch.randint(0, n_src, (e,), generator=gen)
    dst = torch.randint(0, n_dst, (e,), generator=gen)
    if hub:
        dst[: e // 2] = 2
    dst[:5] = src[:5] % n_dst
    ei = torch.stack([src, dst])
    xs = torch.randn(n_src, fs, generator=gen)
    xd = xs if same else torch.randn(n_dst, fd, generator=gen)
    orc, prod = _pair(C, xs, xd, ei, same)
    xs_o = xs.double().requires_grad_(True)
    xd_o = xs_o if same else xd.double().requires_grad_(True)
    xs_p = xs.clone().requires_grad_(True)
    xd_p = xs_p if same else xd.clone().requires_grad_(True)
    w = torch.randn(n_dst, C, generator=gen)
    out_o = orc(xs_o if same else (xs_o, xd_o), ei)
    (out_o * w.double()).sum().backward()
    with cpu_ops():
        agx.functional.GATPlan._cache.clear()
        out_p = prod(xs_p if same else (xs_p, xd_p), ei)
        (out_p * w).sum().backward()
    assert rel_err(out_p, out_o) <= 1e-5
    assert rel_err(xs_p.grad, xs_o.grad) <= 5e-5
    if not same:
        assert rel_err(xd_p.grad, xd_o.grad) <= 5e-5
    po, pp = dict(orc.named_parameters()), dict(prod.named_parameters())
    for k in po:
        assert pp[k].grad.shape == pp[k].shape
        assert rel_err(pp[k].grad, po[k].grad) <= 5e-5, k


class _OneConv(torch.nn.Module):
    def __init__(self, op, C):
        super().__init__()
        self.conv = op((-1, -1), C)

    def forward(self, x, edge_index):
        return self.conv(x, edge_index)


def test_hetero_gat_layer_host_logic_vs_oracle():
    """Three node types (one with one-hot features), five relations: two destinations with several
    incoming relations (short rows, hub rows, both), a homogeneous relation, a type that is only a
    source.  Outputs and all gradients against the oracle's per-relation GATConv + relation sum."""
    gen = torch.Generator().manual_seed(5)
    n = {'a': 300, 'b': 7, 'c': 40}
    feats = {'a': torch.randn(300, 20, generator=gen), 'b': torch.eye(7),
             'c': torch.randn(40, 12, generator=gen)}
    C = 16

    def edges(ns, nd, e, hub=None):
        s = torch.randint(0, ns, (e,), generator=gen)
        d = torch.randint(0, nd, (e,), generator=gen)
        if hub is not None:
            d[: e // 2] = hub
        return torch.stack([s, d])
    ei = OrderedDict([
        (('a', 'r1', 'b'), edges(300, 7, 3000, hub=3)),     # hub rows (avg 430, one > 1024)
        (('c', 'r2', 'b'), edges(40, 7, 30)),               # short rows into the same type
        (('b', 'r3', 'a'), edges(7, 300, 500)),             # hub SOURCES (transpose is long)
        (('c', 'r4', 'a'), edges(40, 300, 600)),
        (('a', 'r5', 'a'), edges(300, 300, 900)),           # homogeneous
    ])
    md = (list(n.keys()), list(ei.keys()))
    convs_o = torch.nn.ModuleDict()
    for (s, r, d) in ei:
        cv = go.GATConv((-1, -1), C)
        with torch.no_grad():
            cv((feats[s], feats[d]), ei[(s, r, d)])
        convs_o['__'.join((s, r, d))] = cv
    util.fill_params_deterministic(convs_o)
    with torch.no_grad():
        for i, cv in enumerate(convs_o.values()):
            cv.bias.copy_(torch.linspace(-0.2, 0.3, C) * (i + 1))
    prod = agx.to_hetero(_OneConv(agx.GATConv, C), md, aggr='sum')
    util.copy_state(convs_o, prod.conv)
    convs_o = convs_o.double()
    x_o = {t: v.double().requires_grad_(t != 'b') for t, v in feats.items()}
    x_p = {t: v.clone().requires_grad_(t != 'b') for t, v in feats.items()}
    outs_o = {}
    for (s, r, d), e in ei.items():
        o = convs_o['__'.join((s, r, d))]((x_o[s], x_o[d]), e)
        outs_o[d] = o if d not in outs_o else outs_o[d] + o
    w = {t: torch.randn(n[t], C, generator=gen) for t in outs_o}
    sum((outs_o[t] * w[t].double()).sum() for t in outs_o).backward()
    with cpu_ops():
        agx.functional.GATPlan._cache.clear()
        outs_p = prod(x_p, ei)
        assert list(outs_p.keys()) == list(outs_o.keys())
        sum((outs_p[t] * w[t]).sum() for t in outs_p).backward()
    for t in outs_o:
        assert rel_err(outs_p[t], outs_o[t]) <= 1e-5, t
    for t in ('a', 'c'):
        assert rel_err(x_p[t].grad, x_o[t].grad) <= 5e-5, t
    po, pp = dict(convs_o.named_parameters()), dict(prod.conv.named_parameters())
    assert set(po) == set(pp)
    gmax = max(float(p.grad.abs().max()) for p in po.values())
    for k in po:
        err = float((pp[k].grad.double() - po[k].grad).abs().max())
        assert err <= 5e-5 * float(po[k].grad.abs().max()) or err <= 1e-7 * gmax, k

    # only some destination types receive a gradient (conv_out: only 'artwork' feeds the loss)
    for p in pp.values():
        p.grad = None
    with cpu_ops():
        outs_p = prod(x_p, ei)
        (outs_p['b'] * w['b']).sum().backward()
    for k, p in pp.items():
        live = k.split('.')[0].endswith('__b')
        assert (p.grad is not None) == live, k


@pytest.mark.parametrize('opname', ['SAGEConv', 'GraphConv'])
def test_hetero_sage_graphconv_layer_host_logic_vs_oracle(opname):
    """The same check for the fused SAGEConv / GraphConv layer (functional._HeteroConvFn):
    transform-first and aggregate-first relations, one-hot inputs, hub rows in both directions."""
    gen = torch.Generator().manual_seed(9)
    n = {'a': 300, 'b': 7, 'c': 40}
    feats = {'a': torch.randn(300, 20, generator=gen), 'b': torch.eye(7),
             'c': torch.randn(40, 12, generator=gen)}
    C = 16

    def edges(ns, nd, e, hub=None):
        s = torch.randint(0, ns, (e,), generator=gen)
        d = torch.randint(0, nd, (e,), generator=gen)
        if hub is not None:
            d[: e // 2] = hub
        return torch.stack([s, d])
    ei = OrderedDict([
        (('a', 'r1', 'b'), edges(300, 7, 3000, hub=3)),
        (('c', 'r2', 'b'), edges(40, 7, 30)),
        (('b', 'r3', 'a'), edges(7, 300, 500)),
        (('c', 'r4', 'a'), edges(40, 300, 600)),
        (('a', 'r5', 'a'), edges(300, 300, 900)),
    ])
    md = (list(n.keys()), list(ei.keys()))
    convs_o = torch.nn.ModuleDict()
    for (s, r, d) in ei:
        cv = getattr(go, opname)((-1, -1), C)
        with torch.no_grad():
            cv((feats[s], feats[d]), ei[(s, r, d)])
        convs_o['__'.join((s, r, d))] = cv
    util.fill_params_deterministic(convs_o)
    prod = agx.to_hetero(_OneConv(getattr(agx, opname), C), md, aggr='sum')
    util.copy_state(convs_o, prod.conv)
    convs_o = convs_o.double()
    x_o = {t: v.double().requires_grad_(t != 'b') for t, v in feats.items()}
    x_p = {t: v.clone().requires_grad_(t != 'b') for t, v in feats.items()}
    outs_o = {}
    for (s, r, d), e in ei.items():
        o = convs_o['__'.join((s, r, d))]((x_o[s], x_o[d]), e)
        outs_o[d] = o if d not in outs_o else outs_o[d] + o
    w = {t: torch.randn(n[t], C, generator=gen) for t in outs_o}
    sum((outs_o[t] * w[t].double()).sum() for t in outs_o).backward()
    with cpu_ops():
        agx.graph.clear_plan_cache()
        outs_p = prod(x_p, ei)
        assert list(outs_p.keys()) == list(outs_o.keys())
        sum((outs_p[t] * w[t]).sum() for t in outs_p).backward()
        agx.graph.clear_plan_cache()
    for t in outs_o:
        assert rel_err(outs_p[t], outs_o[t]) <= 1e-5, t
    for t in ('a', 'c'):
        assert rel_err(x_p[t].grad, x_o[t].grad) <= 5e-5, t
    po, pp = dict(convs_o.named_parameters()), dict(prod.conv.named_parameters())
    assert set(po) == set(pp)
    for k in po:
        assert rel_err(pp[k].grad, po[k].grad) <= 5e-5, k


@pytest.mark.parametrize('opname', ['SAGEConv', 'GraphConv', 'GATConv'])
def test_whole_model_host_logic_vs_oracle(opname):
    """HeteroSGNN through to_hetero on the tiny ArtGraph (one-hot inputs, 17 relations, BatchNorm,
    fused ReLU + dropout with injected masks, log_softmax, nll_loss): the product's host logic
    (tracing, fusion plan, descriptor wiring, gradient delivery) against the oracle, every agx
    entry point restated in torch (tests/cpu_shim.py)."""
    import copy
    g, ei, md = util.undirected_graph('tiny')
    C = 32
    orc = go.HeteroSGNNOracle(getattr(go, opname), torch.nn.ReLU(), 'sum', 128, C, md, 2, 0.4,
                              True, False)
    with torch.no_grad():
        orc(g.x_dict, ei)
    util.fill_params_deterministic(orc)
    util.reset_bn(orc)
    prod = agx.HeteroSGNN(getattr(agx, opname), torch.nn.ReLU(), 'sum', 128, C, md, 2, 0.4, True,
                          False)
    util.copy_state(orc, prod)
    gen = torch.Generator().manual_seed(77)
    masks = {t: (torch.rand(n, 128, generator=gen) >= 0.4).float() / 0.6
             for t, n in g.num_nodes_dict.items()}
    o64 = copy.deepcopy(orc).double()
    o64.gnn.dropout_masks = {t: m.double() for t, m in masks.items()}
    prod.gnn.dropout_masks = masks
    o64.train(); prod.train()
    y = g['artwork'].y_style
    e_o, o_o = o64({k: v.double() for k, v in g.x_dict.items()}, ei)
    l_o = go.nll_loss_artwork(o_o[0], y)
    l_o.backward()
    with cpu_ops():
        agx.graph.clear_plan_cache()
        agx.functional.GATPlan._cache.clear()
        e_p, o_p = prod(g.x_dict, ei)
        l_p = agx.functional.nll_loss(o_p[0]['artwork'], y)
        l_p.backward()
        agx.graph.clear_plan_cache()
        agx.functional.GATPlan._cache.clear()
    assert list(e_p.keys()) == list(e_o.keys())
    for t in e_o:
        assert rel_err(e_p[t], e_o[t]) <= 2e-5, ('emb', t)
        assert rel_err(o_p[0][t], o_o[0][t]) <= 2e-5, ('logp', t)
    assert rel_err(l_p, l_o) <= 1e-5
    og = {n: p.grad for n, p in o64.named_parameters() if p.grad is not None}
    pg = {n: p.grad for n, p in prod.named_parameters() if p.grad is not None}
    if opname == 'GraphConv':
        og = {n.replace('.lin_l.', '.lin_rel.').replace('.lin_r.', '.lin_root.'): v
              for n, v in og.items()}
    gmax = max(float(v.abs().max()) for v in og.values())
    for n, v in og.items():
        if float(v.abs().max()) == 0.0 and n not in pg:
            continue
        err = float((pg[n].double() - v).abs().max())
        # (noise-level tensors -- biases in front of a training-mode BatchNorm have a zero true
        # gradient -- are compared on the scale of the model's gradients, as in the GPU suite)
        assert err <= 1e-4 * max(float(v.abs().max()), 1e-2 * gmax), (n, err)
    for (n, b_o), (_, b_p) in zip(o64.named_buffers(), prod.named_buffers()):
        assert rel_err(b_p, b_o) <= 1e-5, n


def test_identity_marker_matches_dense_one_hot_on_cpu_shim():
    """data.Identity(n) in x_dict: same layer plan (transposed weight copies for one-hot sources)
    and same numbers as a dense torch.eye(n) -- host-side wiring only, kernels restated in torch."""
    import copy
    from collections import OrderedDict
    import mmac_b200 as agx
    from cpu_shim import cpu_ops
    g, ei, md = util.undirected_graph('tiny')
    torch.manual_seed(3)
    m1 = agx.HeteroSGNN(agx.SAGEConv, torch.nn.ReLU(), 'sum', 128, 32, md, 2, 0.0, True, False)
    with cpu_ops(), torch.no_grad():
        m1(g.x_dict, ei)                      # materialise the lazy weights before copying
    m2 = copy.deepcopy(m1)
    xm = OrderedDict((k, v if k == 'artwork' else agx.Identity(v.shape[0]))
                     for k, v in g.x_dict.items())
    with cpu_ops():
        e1, o1 = m1(g.x_dict, ei)
        e2, o2 = m2(xm, ei)
    assert torch.equal(e1['artwork'], e2['artwork'])
    assert torch.equal(o1[0]['artwork'], o2[0]['artwork'])


def test_batch_norm_without_mask_tensor_host_logic():
    """BNSpec.drop (dropout decided inside the normalising pass, y not written, backward gate read
    from the activation and scaled by 1 / (1 - p)) against the same layer with the materialised
    mask: identical activations and gradients.  (Kernel layer restated on the CPU; the stand-in
    dropout stream of tests/cpu_shim.py is a function of (key, counter) like the Philox one.)"""
    import cpu_shim
    import mmac_b200.functional as AF
    gen = torch.Generator().manual_seed(5)
    sizes, F, p = [40, 7], 16, 0.25
    xs = [torch.randn(n, F, generator=gen) for n in sizes]
    gas = [torch.randn(n, F, generator=gen) for n in sizes]
    seed = torch.tensor([12345, 77], dtype=torch.int64)
    offs = [0, ((sizes[0] * F + 3) // 4) * 4]
    masks = [cpu_shim._dropout_mask((n, F), p, [int(seed[0]), int(seed[1]) + o // 4])
             for n, o in zip(sizes, offs)]

    def run(virtual):
        bns = [torch.nn.BatchNorm1d(F).train() for _ in sizes]
        xd = [x.clone().requires_grad_(True) for x in xs]
        spec = AF.BNSpec(n=len(sizes), F=F, training=True, momentum=0.1, eps=1e-5,
                         running=[(b.running_mean, b.running_var) for b in bns], with_act=True,
                         dmasks=None if virtual else masks,
                         drop=[(seed, o) for o in offs] if virtual else None, drop_p=p,
                         need_y=not virtual)
        with cpu_ops():
            res = AF.batch_norm_act(spec, xd, [b.weight for b in bns], [b.bias for b in bns])
            assert len(res) == (len(sizes) if virtual else 2 * len(sizes))
            acts = res[-len(sizes):]
            sum((a * g).sum() for a, g in zip(acts, gas)).backward()
        return acts, [x.grad for x in xd], [b.weight.grad for b in bns], [b.bias.grad for b in bns]

    a0, dx0, dw0, db0 = run(False)
    a1, dx1, dw1, db1 = run(True)
    for i in range(len(sizes)):
        assert torch.equal(a0[i], a1[i])
        assert 0 < int((a1[i] == 0).sum()) < a1[i].numel()
        assert torch.allclose(dx0[i], dx1[i], rtol=1e-6, atol=1e-7)
        assert torch.allclose(dw0[i], dw1[i], rtol=1e-6, atol=1e-7)
        assert torch.allclose(db0[i], db1[i], rtol=1e-6, atol=1e-7)


def test_nll_of_log_softmax_fused_backward_host_logic():
    """functional.nll_loss on the tensor functional.log_softmax returned differentiates both in one
    call (g * (softmax - onehot) on the logits); on any other tensor (a slice, a clone) it takes the
    general two-step path.  Both equal torch's F.nll_loss(F.log_softmax(x))."""
    import mmac_b200.functional as AF
    gen = torch.Generator().manual_seed(6)
    x = torch.randn(37, 12, generator=gen)
    y = torch.randint(0, 12, (37,), generator=gen)
    xr = x.clone().requires_grad_(True)
    ref = torch.nn.functional.nll_loss(torch.log_softmax(xr, 1), y)
    ref.backward()
    for fused in (True, False):
        xd = x.clone().requires_grad_(True)
        with cpu_ops():
            lp = AF.log_softmax(xd, 1)
            assert getattr(lp, '_agx_logits', None) is xd
            arg = lp if fused else lp.clone()
            loss = AF.nll_loss(arg, y)
            assert (type(loss.grad_fn).__name__ == '_NLLOfLogSoftmaxFnBackward') == fused
            loss.backward()
        assert torch.allclose(loss, ref, rtol=1e-6)
        assert torch.allclose(xd.grad, xr.grad, rtol=1e-5, atol=1e-8)

"""CPU suite, part 1: the oracle against the committed golden fixtures (made by the reference's own
model code, tests/golden/make_golden.py) and against independent restatements."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import util
from util import rel_err
import mmac_b200  # noqa: F401
from mmac_b200 import synth
from oracle import graph_oracle as go
from oracle import heads_oracle as ho


def _oracle_gnn(op, label, size='tiny', dtype=torch.float32):
    g, ei, md = util.undirected_graph(size)
    C = {'style': 32, 'genre': 18}[label]
    m = go.HeteroSGNNOracle(op, torch.nn.ReLU(), 'sum', 128, C, md, 2, 0.0, True, False)
    with torch.no_grad():
        m(g.x_dict, ei)
    util.fill_params_deterministic(m)
    util.reset_bn(m)
    return g, ei, m


@pytest.mark.parametrize('opname,label', [('SAGEConv', 'style'), ('GraphConv', 'genre'),
                                          ('GATConv', 'style')])
def test_oracle_matches_reference_wiring_golden(opname, label):
    gold = util.load_golden(f'gnn_tiny_{opname.lower()}_{label}.npz')
    op = getattr(go, opname)
    g, ei, m = _oracle_gnn(op, label)
    m.train()
    emb, out = m(g.x_dict, ei)
    loss = go.nll_loss_artwork(out[0], g['artwork'][f'y_{label}'])
    loss.backward()
    # same ATen kernels, same order of operations: bit-exact
    assert np.array_equal(emb['artwork'].detach().numpy(), gold['emb_artwork'])
    assert np.array_equal(emb['style'].detach().numpy(), gold['emb_style'])
    assert np.array_equal(emb['genre'].detach().numpy(), gold['emb_genre'])
    assert np.array_equal(out[0]['artwork'].detach().numpy(), gold['logp_artwork'])
    assert np.array_equal(out[0]['tag'].detach().numpy(), gold['logp_tag'])
    assert np.float32(loss.item()) == gold['loss']
    sd = m.state_dict()
    assert np.array_equal(sd['gnn.bns.1.artwork.running_mean'].numpy(), gold['running_mean_bn1_artwork'])
    assert np.array_equal(sd['gnn.bns.1.artwork.running_var'].numpy(), gold['running_var_bn1_artwork'])
    assert np.array_equal(ei[('artist', 'teacher_rel', 'artist')].numpy(), gold['teacher_edge_index'])
    named = dict(m.named_parameters())
    for k, v in gold.items():
        if k.startswith('grad::'):
            name = k[6:]
            if opname == 'GraphConv':
                name = name.replace('.lin_l.', '.lin_rel.').replace('.lin_r.', '.lin_root.')
            assert rel_err(named[name].grad, v) <= 1e-6, k


def test_to_undirected_semantics():
    g = synth.make_artgraph('tiny')
    ei = go.to_undirected(g.edge_index_dict)
    keys = list(ei.keys())
    assert len(keys) == 17
    assert keys[:9] == [tuple(k) for k in synth.EDGE_TYPES]
    assert [k[1] for k in keys[9:]] == ['rev_field_rel', 'rev_movement_rel', 'rev_media_rel',
                                        'rev_about_rel', 'rev_genre_rel', 'rev_style_rel',
                                        'rev_author_rel', 'rev_locatedin_rel']
    for (s, r, d), v in g.edge_index_dict.items():
        if s != d:
            rv = ei[(d, 'rev_' + r, s)]
            assert torch.equal(rv[0], v[1]) and torch.equal(rv[1], v[0])   # order preserved
    t = ei[('artist', 'teacher_rel', 'artist')]
    n = g.num_nodes_dict['artist']
    key = t[0] * n + t[1]
    assert torch.all(key[1:] > key[:-1])                                  # sorted, unique
    pairs = set(map(tuple, t.t().tolist()))
    assert all((b, a) in pairs for a, b in pairs)                         # symmetric
    raw = g.edge_index_dict[('artist', 'teacher_rel', 'artist')]
    assert all((int(a), int(b)) in pairs for a, b in raw.t().tolist())


@pytest.mark.parametrize('reduce', ['mean', 'add'])
def test_propagate_vs_scipy_spmm(reduce):
    gen = torch.Generator().manual_seed(3)
    n_src, n_dst, e, f = 57, 23, 400, 24
    ei = torch.stack([torch.randint(0, n_src, (e,), generator=gen),
                      torch.randint(0, n_dst - 3, (e,), generator=gen)])      # last rows isolated
    x = torch.randn(n_src, f, generator=gen, dtype=torch.float64)
    out = go.propagate(x, ei, n_dst, reduce)
    A = sp.coo_matrix((np.ones(e), (ei[1].numpy(), ei[0].numpy())), shape=(n_dst, n_src)).tocsr()
    ref = A @ x.numpy()
    if reduce == 'mean':
        deg = np.maximum(np.asarray(A.sum(axis=1)).ravel(), 1)
        ref = ref / deg[:, None]
    assert np.allclose(out.numpy(), ref, rtol=1e-12, atol=1e-12)
    assert torch.all(out[-3:] == 0)


def test_csr_oracle_is_stable_sort():
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 17, size=300)
    vals = rng.integers(0, 1000, size=300)
    rowptr, col, eid = go.csr_build(keys, vals, 20)
    assert rowptr[0] == 0 and rowptr[-1] == 300 and np.all(np.diff(rowptr) >= 0)
    for r in range(20):
        idx = np.nonzero(keys == r)[0]                     # edge-list order
        assert np.array_equal(eid[rowptr[r]:rowptr[r + 1]], idx)
        assert np.array_equal(col[rowptr[r]:rowptr[r + 1]], vals[idx])
    t_sorted, t_perm = torch.sort(torch.from_numpy(keys), stable=True)
    assert np.array_equal(eid, t_perm.numpy())


def test_scatter_add_accumulates_in_edge_order():
    """The claim K2's determinism rests on: CPU scatter_add_ == sequential
    edge-order accumulation, bit for bit."""
    gen = torch.Generator().manual_seed(5)
    e, n_dst, f = 5000, 7, 16
    ei = torch.stack([torch.arange(e), torch.randint(0, n_dst, (e,), generator=gen)])
    x = torch.randn(e, f, generator=gen) * 100
    out = go.propagate(x, ei, n_dst, 'add')
    seq = torch.zeros(n_dst, f)
    for k in range(e):
        seq[ei[1, k]] += x[k]
    assert torch.equal(out, seq)


def test_transform_first_equals_aggregate_first():
    gen = torch.Generator().manual_seed(9)
    n_src, n_dst, e = 40, 300, 900
    ei = torch.stack([torch.randint(0, n_src, (e,), generator=gen),
                      torch.randint(0, n_dst, (e,), generator=gen)])
    x = torch.randn(n_src, 64, generator=gen)
    w = torch.randn(32, 64, generator=gen) / 8
    a = go.propagate(x, ei, n_dst, 'mean') @ w.t()
    b = go.propagate(x @ w.t(), ei, n_dst, 'mean')
    assert rel_err(b, a) < 1e-6


def test_fp32_oracle_close_to_fp64():
    g, ei, m = _oracle_gnn(go.SAGEConv, 'style')
    m.eval()
    m.gnn.traced_training = False
    with torch.no_grad():
        e32, o32 = m(g.x_dict, ei)
    m64 = m.double()
    with torch.no_grad():
        e64, o64 = m64({k: v.double() for k, v in g.x_dict.items()}, ei)
    assert rel_err(e32['artwork'], e64['artwork']) < 1e-5
    assert rel_err(o32[0]['artwork'], o64[0]['artwork']) < 1e-5


@pytest.mark.parametrize('arch,fv', [('vit', 768), ('resnet', 2048)])
def test_heads_oracle_matches_reference_golden(arch, fv):
    gold = util.load_golden(f'heads_{arch}.npz')
    n = 96
    feat, emb_s, emb_g, y_s, y_g = synth.make_head_batch(n, arch=arch, seed=7)
    m = ho.MultiTaskHeadOracle(fv, 128, {'style': 32, 'genre': 18}, 0.0)
    util.fill_params_deterministic(m)
    f = feat.clone().requires_grad_(True)
    out = m(f, emb_s, emb_g)
    w_s, w_g = ho.class_weights(y_s, 32), ho.class_weights(y_g, 18)
    loss = ho.multitask_loss(out, y_s, y_g, w_s, w_g)
    loss.backward()
    assert np.array_equal(out[0].detach().numpy(), gold['out_style'])
    assert np.array_equal(out[1].detach().numpy(), gold['out_genre'])
    assert np.float32(loss.item()) == gold['loss_weighted']
    assert rel_err(m.class_style[1].weight.grad, gold['grad_w_style']) <= 1e-6
    assert rel_err(m.class_genre[1].bias.grad, gold['grad_b_genre']) <= 1e-6
    assert rel_err(f.grad, gold['grad_feat']) <= 1e-6
    lu = ho.multitask_loss([o.detach() for o in out], y_s, y_g)
    assert np.float32(lu.item()) == gold['loss_unweighted']

    m1 = ho.SingleTaskHeadOracle(fv, 128, 32, 0.0)
    util.fill_params_deterministic(m1)
    assert np.array_equal(m1(feat, emb_s).detach().numpy(), gold['single_out'])

    mp = ho.ProjectorOracle(fv, 128)
    util.fill_params_deterministic(mp)
    o = mp(feat)
    lp = ho.projector_loss(o, emb_s * 3.0)
    lp.backward()
    assert np.array_equal(o.detach().numpy(), gold['proj_out'])
    assert np.float32(lp.item()) == gold['proj_loss']
    assert rel_err(mp.encoder.weight.grad, gold['proj_grad_w']) <= 1e-6


def test_class_weights_formula():
    y = torch.tensor([0, 0, 0, 1, 2, 2])
    w = ho.class_weights(y, 3)
    assert torch.allclose(w, torch.tensor([6 / (3 * 3), 6 / (1 * 3), 6 / (2 * 3)]))
    assert torch.allclose(synth.class_weights(y, 3), w)


def test_synth_graph_schema():
    g = synth.make_artgraph('small')
    assert g.node_types == synth.NODE_TYPES
    assert g.edge_types == [tuple(e) for e in synth.EDGE_TYPES]
    assert g['artwork'].x.shape == (2000, 128) and g['artwork'].x.dtype == torch.float32
    assert torch.equal(g['style'].x, torch.eye(32))
    assert g['artwork'].y_style.dtype == torch.float32
    n = g.num_nodes_dict
    for (s, r, d), ei in g.edge_index_dict.items():
        assert ei.dtype == torch.int64 and ei.shape[0] == 2 and ei.is_contiguous()
        assert int(ei[0].max()) < n[s] and int(ei[1].max()) < n[d]
    st = g[('artwork', 'style_rel', 'style')].edge_index
    order = torch.argsort(st[0])
    assert torch.equal(st[1][order].float(), g['artwork'].y_style)         # labels = edge targets
    g2 = synth.make_artgraph('small')
    assert all(torch.equal(a, b) for a, b in zip(g.edge_index_dict.values(),
                                                 g2.edge_index_dict.values()))
    r = synth.replicate(synth.make_artgraph('tiny'), 3)
    assert r['artwork'].x.shape[0] == 900 and r.num_edges() == 3 * synth.make_artgraph('tiny').num_edges()


def test_context_heads_oracle_matches_reference_golden():
    """ContextNet / Castellano heads: the oracle classes reproduce the
    outputs, combined losses and gradients of the reference's own classes (heads_context.npz)."""
    gold = util.load_golden('heads_context.npz')
    n = 64
    feat, emb_s, _, y_s, y_g = synth.make_head_batch(n, arch='resnet', seed=11)

    m = ho.ContextNetSingleOracle(2048, 128, 32)
    util.fill_params_deterministic(m)
    out, proj = m(feat)
    loss = ho.context_loss(out, proj, y_s, emb_s * 3.0, 0.9, 'smooth_l1')
    loss.backward()
    assert np.array_equal(out.detach().numpy(), gold['cn1_out'])
    assert np.array_equal(proj.detach().numpy(), gold['cn1_proj'])
    assert abs(loss.item() - float(gold['cn1_loss'])) <= 1e-6 * abs(float(gold['cn1_loss']))
    assert np.allclose(m.classifier.weight.grad.numpy(), gold['cn1_grad_cls_w'], rtol=1e-5, atol=1e-9)
    assert np.allclose(m.encoder.weight.grad.numpy()[:8], gold['cn1_grad_enc_w'], rtol=1e-5, atol=1e-9)

    m = ho.ContextNetMultiOracle(2048, 128, {'style': 32, 'genre': 18})
    util.fill_params_deterministic(m)
    outs, proj = m(feat)
    loss = ho.context_loss(outs, proj, (y_s, y_g), emb_s * 3.0, 0.9, 'smooth_l1')
    loss.backward()
    assert np.array_equal(outs[0].detach().numpy(), gold['cn2_out_style'])
    assert np.array_equal(outs[1].detach().numpy(), gold['cn2_out_genre'])
    assert abs(loss.item() - float(gold['cn2_loss'])) <= 1e-6 * abs(float(gold['cn2_loss']))
    assert np.allclose(m.encoder.bias.grad.numpy(), gold['cn2_grad_enc_b'], rtol=1e-5, atol=1e-9)

    for tag, m in (('mm1', ho.CastellanoSingleOracle(2048, 128, 32, 0.0)),
                   ('mm2', ho.CastellanoMultiOracle(2048, 128, {'style': 32, 'genre': 18}, 0.0))):
        util.fill_params_deterministic(m)
        f_in = feat.clone().requires_grad_(True)
        out, proj = m(f_in)
        labels = y_s if tag == 'mm1' else (y_s, y_g)
        loss = ho.context_loss(out, proj, labels, emb_s, 0.6, 'mse')
        loss.backward()
        assert np.array_equal(proj.detach().numpy(), gold[f'{tag}_proj'])
        assert abs(loss.item() - float(gold[f'{tag}_loss'])) <= 1e-6 * abs(float(gold[f'{tag}_loss']))
        assert np.allclose(m.encoder[0].weight.grad.numpy()[:8], gold[f'{tag}_grad_enc0_w'],
                           rtol=1e-5, atol=1e-9)
        assert np.allclose(f_in.grad.numpy()[:8], gold[f'{tag}_grad_feat'], rtol=1e-5, atol=1e-9)


def test_gatconv_oracle_vs_dense_attention():
    """Independent second opinion for the GATConv restatement: the same layer written as dense
    masked attention (an [N_dst, N_src] multiplicity matrix, torch.softmax over the rows) in
    float64 -- duplicate edges count twice, existing self loops are dropped and (i, i) added for
    i < min(N_src, N_dst), rows without any edge give the bias."""
    gen = torch.Generator().manual_seed(17)
    n_src, n_dst, C = 23, 31, 8
    ei = torch.stack([torch.randint(0, n_src, (90,), generator=gen),
                      torch.randint(0, n_dst - 3, (90,), generator=gen)])
    ei[:, :4] = torch.tensor([[1, 1, 5, 5], [1, 1, 7, 7]])       # a repeated self loop, a duplicate
    xs, xd = torch.randn(n_src, 6, generator=gen), torch.randn(n_dst, 9, generator=gen)
    conv = go.GATConv((-1, -1), C)
    with torch.no_grad():
        conv((xs, xd), ei)
    util.fill_params_deterministic(conv)
    with torch.no_grad():
        conv.bias.copy_(torch.linspace(-1, 1, C))
    conv = conv.double()
    out = conv((xs.double(), xd.double()), ei)

    x_l = xs.double() @ conv.lin_l.weight.t()
    a_l = x_l @ conv.att_l.view(-1)
    a_r = (xd.double() @ conv.lin_r.weight.t()) @ conv.att_r.view(-1)
    mult = torch.zeros(n_dst, n_src, dtype=torch.float64)
    for s, d in ei.t().tolist():
        if s != d:
            mult[d, s] += 1
    for i in range(min(n_src, n_dst)):
        mult[i, i] += 1
    logits = torch.nn.functional.leaky_relu(a_r[:, None] + a_l[None, :], 0.2)
    w = mult * torch.exp(logits - logits.max())
    alpha = w / w.sum(1, keepdim=True).clamp(min=1e-300)
    want = alpha @ x_l + conv.bias
    assert mult[n_dst - 1].sum() == 0                            # a row with no edge at all
    assert rel_err(out, want) <= 1e-12
    assert torch.allclose(out[n_dst - 1], conv.bias)

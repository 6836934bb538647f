"""The reference's OWN files, unmodified, on this repo's operators (row b-2 of SURVEY.md section 8:
"drops into ... the train_*.py scripts unchanged").

``/root/reference/src/train_gnn_embeddings.py`` is executed as ``__main__`` with
``compat/`` on ``sys.path``: its ``import torch_geometric.nn as operators`` /
``torch_geometric.transforms`` / ``torch_geometric.data`` resolve to ``compat/torch_geometric``
(the product's classes), its ``ArtGraph`` dataset class (``src/data/artgraph.py``, unmodified) reads
a synthetic raw CSV tree written by ``synth.write_artgraph_raw``, its ``models_graph.HeteroSGNN``
(unmodified) is built from the product's ``SAGEConv`` / ``to_hetero``.

/root/reference exists in the build container only and this container has no GPU, so the run uses
the TEST-ONLY torch restatement of the kernel layer (tests/cpu_shim.py): what is proven here is the
interface (imports, constructor / call signatures, dict structures, host tensors in and out,
``torch.optim.Adam`` created before the lazy weights exist, state-dict layout, deepcopy) and that
the trained model equals the oracle with the same weights.  The kernels themselves are checked on
the GPU by tests/test_gpu_*.py (``test_host_io_*`` covers this calling convention there).
"""
import os
import runpy
import sys

import pytest
import torch

import util  # noqa: F401
from cpu_shim import cpu_ops
from oracle import graph_oracle as go

REF_SRC = '/root/reference/src'
COMPAT = os.path.join(util.ROOT, 'compat')

pytestmark = pytest.mark.skipif(not os.path.isdir(REF_SRC),
                                reason='the reference tree is only present in the build container')


def _make_dataset_tree(tmp):
    from mmac_b200 import synth
    graphs = {}
    for i, split in enumerate(('train', 'train_train', 'train_validation', 'train_test')):
        g = synth.make_artgraph('tiny', features='one-hot', seed=900 + i)
        synth.write_artgraph_raw(os.path.join(tmp, 'dataset', split), g)
        graphs[split] = g
    os.makedirs(os.path.join(tmp, 'dataset', 'train', 'embeddings'))
    os.makedirs(os.path.join(tmp, 'src'))
    return graphs


@pytest.mark.parametrize('operator', ['SAGEConv', 'GraphConv'])
def test_reference_train_gnn_embeddings_script_runs_unchanged(tmp_path, monkeypatch, operator):
    tmp = str(tmp_path)
    graphs = _make_dataset_tree(tmp)
    monkeypatch.chdir(os.path.join(tmp, 'src'))           # config.py paths are relative ('../dataset')
    monkeypatch.setattr(sys, 'argv', ['train_gnn_embeddings.py', '--label', 'style',
                                      '--operator', operator, '--epochs', '6'])
    monkeypatch.syspath_prepend(REF_SRC)
    monkeypatch.syspath_prepend(COMPAT)
    for name in [m for m in sys.modules if m.split('.')[0] in
                 ('torch_geometric', 'models', 'data', 'config')]:
        monkeypatch.delitem(sys.modules, name)
    torch.manual_seed(0)
    with cpu_ops():
        ns = runpy.run_path(os.path.join(REF_SRC, 'train_gnn_embeddings.py'), run_name='__main__')
        model = ns['model']
        import torch_geometric
        assert torch_geometric.__version__.endswith('+agx')
        import mmac_b200 as agx
        assert isinstance(model.gnn, agx.HeteroModule) and model.gnn.host_io
        assert type(model).__module__ == 'models.models_graph'       # the reference's class
        # the dataset class read the synthetic CSV tree into the reference's schema
        data_train = ns['data_train']
        g = graphs['train_train']
        assert torch.equal(data_train['artwork'].x, g['artwork'].x)
        assert data_train['tag'].x.shape == (40, 40) and len(data_train.metadata()[1]) == 17
        assert torch.equal(data_train[('artwork', 'style_rel', 'style')].edge_index,
                           g[('artwork', 'style_rel', 'style')].edge_index)
        # state-dict layout of PyG's to_hetero
        sd = model.state_dict()
        root = 'lin_l' if operator == 'SAGEConv' else 'lin_rel'
        assert f'gnn.convs.0.artist__field_rel__field.{root}.weight' in sd
        assert 'gnn.bns.1.artwork.running_mean' in sd
        assert sd[f'gnn.conv_out.tag__rev_about_rel__artwork.{root}.weight'].shape == (32, 128)
        # training ran: hetero_training() returns finite, decreasing losses
        _, losses, accs = ns['hetero_training']()
        first = float(losses[0])
        for _ in range(5):
            _, losses, accs = ns['hetero_training']()
        assert torch.isfinite(losses[0]) and float(losses[0]) < first
        val_l, val_a, test_l, test_a = ns['hetero_test']()
        assert torch.isfinite(val_l[0]) and 0.0 <= float(test_a[0]) <= 1.0
        # save_embeddings(): deepcopy + eval forward on the full-train graph, torch.save
        ns['save_embeddings'](model, 'style')
        emb = torch.load(os.path.join(tmp, 'dataset', 'train', 'embeddings',
                                      'test_gnn_artwork_style_embs.pt'))
        assert emb.shape == (300, 128) and emb.device.type == 'cpu'
        # ... and is what the oracle computes from the same weights on the same graph
        full = ns['data_train_full']
        orc = go.HeteroSGNNOracle(getattr(go, operator), torch.nn.ReLU(), 'sum', 128, 32,
                                  full.metadata(), 2, 0.4, True, False)
        with torch.no_grad():
            orc(full.x_dict, full.edge_index_dict)
        util.copy_state(model, orc)
        orc.eval()
        with torch.no_grad():
            emb_o, _ = orc(full.x_dict, full.edge_index_dict)
    assert util.rel_err(emb, emb_o['artwork']) <= 1e-5


def test_reference_models_graph_on_compat_matches_golden(monkeypatch):
    """``models_graph.HeteroSGNN`` (unmodified) on the compat package reproduces the golden fixture
    that the same file produced on the oracle operators (tests/golden/make_golden.py)."""
    monkeypatch.syspath_prepend(REF_SRC)
    monkeypatch.syspath_prepend(COMPAT)
    for name in [m for m in sys.modules if m.split('.')[0] in ('torch_geometric', 'models')]:
        monkeypatch.delitem(sys.modules, name)
    import torch_geometric.nn as operators
    from models.models_graph import HeteroSGNN
    gold = util.load_golden('gnn_tiny_sageconv_style.npz')
    g, ei, md = util.undirected_graph('tiny')
    orc = go.HeteroSGNNOracle(go.SAGEConv, torch.nn.ReLU(), 'sum', 128, 32, md, 2, 0.0, True, False)
    with torch.no_grad():
        orc(g.x_dict, ei)
    util.fill_params_deterministic(orc)
    util.reset_bn(orc)
    model = HeteroSGNN(operators.SAGEConv, torch.nn.ReLU(), 'sum', 128, 32, md, 2, 0.0, True, False)
    util.copy_state(orc, model)
    model.train()
    with cpu_ops():
        emb, out = model(g.x_dict, ei)
        loss = torch.nn.functional.nll_loss(out[0]['artwork'], g['artwork'].y_style.long())
        loss.backward()
    assert util.rel_err(emb['artwork'], gold['emb_artwork']) <= 1e-5
    assert util.rel_err(out[0]['artwork'], gold['logp_artwork']) <= 1e-5
    assert util.rel_err(loss, gold['loss']) <= 1e-5

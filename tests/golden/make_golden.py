"""Generates the golden fixtures in this directory by running the reference's OWN model code,
imported unmodified from /root/reference (present in the build container only).

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz

GNN fixtures: ``/root/reference/src/models/models_graph.py`` (HeteroGNN / HeteroSGNN) is imported
with ``torch_geometric.nn`` bound to the CPU oracle's operators (PyG itself is absent, see
oracle/__init__.py), so the network WIRING is the reference's, the operator arithmetic the
oracle's restatement of PyG 2.0.2.

Head fixtures: ``/root/reference/src/models/models_kg.py`` classes are instantiated with the
backbone constructors stubbed (``timm.create_model`` / ``torchvision.models.resnet50`` return an
identity feature extractor); everything after the backbone is the reference's first-party torch
arithmetic.

Inputs are regenerated from seeds (``mmac_b200.synth``) and weights from a closed-form pattern
(``tests/util.py``), so the fixtures only store outputs.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import mmac_b200  # noqa: E402,F401
from mmac_b200 import synth  # noqa: E402
from oracle import graph_oracle as go  # noqa: E402
import util  # noqa: E402

REF_SRC = '/root/reference/src'


def _stub_pyg():
    tg = types.ModuleType('torch_geometric')
    tgn = types.ModuleType('torch_geometric.nn')
    tgn.SAGEConv, tgn.GraphConv, tgn.GCNConv = go.SAGEConv, go.GraphConv, go.GraphConv
    tgn.GATConv = go.GATConv
    tgn.Linear, tgn.to_hetero = go.Linear, go.to_hetero
    tg.nn = tgn
    sys.modules['torch_geometric'] = tg
    sys.modules['torch_geometric.nn'] = tgn


class _IdentityBackbone(torch.nn.Module):
    """Stands in for ViT-B/16 / ResNet50: the 'image' IS the feature vector."""

    def __init__(self, feat):
        super().__init__()
        self.head = torch.nn.Linear(feat, 1)        # .head.in_features is read at models_kg.py:223
        self.fc = torch.nn.Linear(feat, 1)          # .fc.in_features is read at models_kg.py:170

    def forward_features(self, x):
        return x

    def forward(self, x):
        return x


def _stub_backbones():
    timm = types.ModuleType('timm')
    timm.create_model = lambda name, pretrained=True: _IdentityBackbone(768)
    sys.modules['timm'] = timm
    import torchvision.models as tvm

    def resnet50(pretrained=True):
        net = _IdentityBackbone(2048)
        # models_kg.py:172 rebuilds nn.Sequential(*list(children())[:-1]): make that an identity
        net.children = lambda: iter([torch.nn.Identity(), torch.nn.Identity()])
        return net
    tvm.resnet50 = resnet50


def gnn_fixture(operator_name, label, size):
    from models.models_graph import HeteroSGNN          # the reference's file
    g = synth.make_artgraph(size)
    ei = go.to_undirected(g.edge_index_dict)
    md = (g.node_types, list(ei.keys()))
    op = {'SAGEConv': go.SAGEConv, 'GraphConv': go.GraphConv, 'GATConv': go.GATConv}[operator_name]
    C = {'style': 32, 'genre': 18}[label]
    torch.manual_seed(0)
    model = HeteroSGNN(op, torch.nn.ReLU(), 'sum', 128, C, md, 2, 0.0, True, False)
    with torch.no_grad():
        model(g.x_dict, ei)                              # materialise lazy weights (:146-147)
    util.fill_params_deterministic(model)
    util.reset_bn(model)
    model.train()
    y = g['artwork'][f'y_{label}']
    emb, out = model(g.x_dict, ei)
    loss = torch.nn.functional.nll_loss(out[0]['artwork'], y.type(torch.LongTensor))
    loss.backward()
    sd = model.state_dict()
    probe = ['gnn.convs.0.tag__rev_about_rel__artwork.lin_l.weight',
             'gnn.convs.1.artwork__style_rel__style.lin_l.weight',
             'gnn.convs.1.artist__teacher_rel__artist.lin_r.weight',
             'gnn.conv_out.style__rev_style_rel__artwork.lin_l.weight',
             'gnn.bns.0.artwork.weight', 'gnn.bns.1.tag.bias']
    named = dict(model.named_parameters())
    out_npz = {
        'emb_artwork': emb['artwork'].detach().numpy(),
        'emb_style': emb['style'].detach().numpy(),
        'emb_genre': emb['genre'].detach().numpy(),
        'logp_artwork': out[0]['artwork'].detach().numpy(),
        'logp_tag': out[0]['tag'].detach().numpy(),
        'loss': np.float32(loss.item()),
        'running_mean_bn1_artwork': sd['gnn.bns.1.artwork.running_mean'].numpy(),
        'running_var_bn1_artwork': sd['gnn.bns.1.artwork.running_var'].numpy(),
        'n_undirected_teacher': np.int64(ei[('artist', 'teacher_rel', 'artist')].shape[1]),
        'teacher_edge_index': ei[('artist', 'teacher_rel', 'artist')].numpy(),
    }
    for k in probe:
        kk = k
        if operator_name == 'GraphConv':        # PyG 2.0.x attribute names of GraphConv
            kk = k.replace('.lin_l.', '.lin_rel.').replace('.lin_r.', '.lin_root.')
        out_npz['grad::' + k] = named[kk].grad.numpy()
    return out_npz


def heads_fixture(arch):
    from models import models_kg                          # the reference's file
    n = 96
    feat, emb_s, emb_g, y_s, y_g = synth.make_head_batch(n, arch=arch, seed=7)
    res = {}
    torch.manual_seed(0)
    cls = models_kg.NewMultiModalMultiTaskViT if arch == 'vit' else models_kg.NewMultiModalMultiTask
    m = cls(emb_size=128, num_classes={'style': 32, 'genre': 18}, dropout=0.0)
    util.fill_params_deterministic(m, only=('class_style', 'class_genre'))
    feat_in = feat.clone().requires_grad_(True)
    img = feat_in if arch == 'vit' else feat_in.view(n, -1, 1, 1)
    out = m(img, emb_s, emb_g)
    w_s = synth.class_weights(y_s, 32)
    w_g = synth.class_weights(y_g, 18)
    loss = 0.5 * torch.nn.CrossEntropyLoss(w_s)(out[0], y_s) + \
        0.5 * torch.nn.CrossEntropyLoss(w_g)(out[1], y_g)
    loss.backward()
    res.update({'out_style': out[0].detach().numpy(), 'out_genre': out[1].detach().numpy(),
                'loss_weighted': np.float32(loss.item()),
                'grad_w_style': m.class_style[1].weight.grad.numpy(),
                'grad_b_genre': m.class_genre[1].bias.grad.numpy(),
                'grad_feat': feat_in.grad.numpy()})
    loss_u = 0.5 * torch.nn.CrossEntropyLoss()(out[0].detach(), y_s) + \
        0.5 * torch.nn.CrossEntropyLoss()(out[1].detach(), y_g)
    res['loss_unweighted'] = np.float32(loss_u.item())

    # single task
    cls1 = models_kg.NewMultiModalSingleTaskVit if arch == 'vit' else models_kg.NewMultiModalSingleTask
    m1 = cls1(emb_size=128, num_class=32, dropout=0.0)
    util.fill_params_deterministic(m1, only=('classifier',))
    o1 = m1(feat if arch == 'vit' else feat.view(n, -1, 1, 1), emb_s)
    res['single_out'] = o1.detach().numpy()

    # projector
    clsp = models_kg.LabelProjectorVit if arch == 'vit' else models_kg.LabelProjector
    mp = clsp(emb_size=128)
    util.fill_params_deterministic(mp, only=('encoder',))
    op = mp(feat if arch == 'vit' else feat.view(n, -1, 1, 1))
    lp = torch.nn.SmoothL1Loss()(op, emb_s * 3.0)         # *3: exercise both Huber branches
    lp.backward()
    res.update({'proj_out': op.detach().numpy(), 'proj_loss': np.float32(lp.item()),
                'proj_grad_w': mp.encoder.weight.grad.numpy()})
    return res


def context_fixture():
    """ContextNet / Castellano heads from the reference's classes with the
    ResNet stubbed by the identity: outputs, the combined loss of train_baseline_context*.py and
    gradients."""
    from models import models_kg
    n = 64
    feat, emb_s, _, y_s, y_g = synth.make_head_batch(n, arch='resnet', seed=11)
    img = feat.view(n, -1, 1, 1)
    res = {}
    # ContextNet single task: SmoothL1, lamb 0.9
    m = models_kg.ContextNetSingleTask(emb_size=128, num_class=32)
    util.fill_params_deterministic(m, only=('classifier', 'encoder'))
    out, proj = m(img)
    loss = 0.9 * torch.nn.CrossEntropyLoss()(out, y_s) + 0.1 * torch.nn.SmoothL1Loss()(proj, emb_s * 3.0)
    loss.backward()
    res.update({'cn1_out': out.detach().numpy(), 'cn1_proj': proj.detach().numpy(),
                'cn1_loss': np.float32(loss.item()),
                'cn1_grad_cls_w': m.classifier.weight.grad.numpy(),
                'cn1_grad_enc_w': m.encoder.weight.grad.numpy()[:8]})
    # ContextNet multitask
    m = models_kg.ContextNetlMultiTask(emb_size=128, num_classes={'style': 32, 'genre': 18})
    util.fill_params_deterministic(m, only=('class_style', 'class_genre', 'encoder'))
    outs, proj = m(img)
    loss = 0.9 * (0.5 * torch.nn.CrossEntropyLoss()(outs[0], y_s) +
                  0.5 * torch.nn.CrossEntropyLoss()(outs[1], y_g)) + \
        0.1 * torch.nn.SmoothL1Loss()(proj, emb_s * 3.0)
    loss.backward()
    res.update({'cn2_out_style': outs[0].detach().numpy(), 'cn2_out_genre': outs[1].detach().numpy(),
                'cn2_loss': np.float32(loss.item()),
                'cn2_grad_enc_b': m.encoder.bias.grad.numpy()})
    # Castellano single / multi: MSE, lamb 0.6, Dropout(0.2) switched off by eval-free p=0 copy
    for tag, cls, kw in (('mm1', models_kg.MultiModalSingleTask, dict(num_class=32)),
                         ('mm2', models_kg.MultiModalMultiTask,
                          dict(num_classes={'style': 32, 'genre': 18}))):
        m = cls(emb_size=128, **kw)
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        util.fill_params_deterministic(m, only=('classifier', 'class_style', 'class_genre', 'encoder'))
        f_in = feat.clone().requires_grad_(True)
        out, proj = m(f_in)
        if tag == 'mm1':
            cl = torch.nn.CrossEntropyLoss()(out, y_s)
            res['mm1_out'] = out.detach().numpy()
        else:
            cl = 0.5 * torch.nn.CrossEntropyLoss()(out[0], y_s) + \
                0.5 * torch.nn.CrossEntropyLoss()(out[1], y_g)
            res['mm2_out_style'] = out[0].detach().numpy()
            res['mm2_out_genre'] = out[1].detach().numpy()
        loss = 0.6 * cl + 0.4 * torch.nn.MSELoss()(proj, emb_s)
        loss.backward()
        res.update({f'{tag}_proj': proj.detach().numpy(), f'{tag}_loss': np.float32(loss.item()),
                    f'{tag}_grad_enc0_w': m.encoder[0].weight.grad.numpy()[:8],
                    f'{tag}_grad_enc2_b': m.encoder[2].bias.grad.numpy(),
                    f'{tag}_grad_feat': f_in.grad.numpy()[:8]})
    return res


def main():
    if not os.path.isdir(REF_SRC):
        raise SystemExit('the reference is not mounted; fixtures can only be regenerated in the '
                         'build container')
    _stub_pyg()
    _stub_backbones()
    sys.path.insert(0, REF_SRC)
    for op, label, size in (('SAGEConv', 'style', 'tiny'), ('GraphConv', 'genre', 'tiny'),
                            ('GATConv', 'style', 'tiny')):
        np.savez_compressed(os.path.join(HERE, f'gnn_{size}_{op.lower()}_{label}.npz'),
                            **gnn_fixture(op, label, size))
    for arch in ('vit', 'resnet'):
        np.savez_compressed(os.path.join(HERE, f'heads_{arch}.npz'), **heads_fixture(arch))
    np.savez_compressed(os.path.join(HERE, 'heads_context.npz'), **context_fixture())
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == '__main__':
    main()

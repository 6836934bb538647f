"""GPU parity of the fused tensor-core head step (agx_head_step: TMA -> bf16 -> tcgen05.mma -> TMEM
epilogue with softmax-CE / SmoothL1 -> second tcgen05.mma pass for the weight gradient) against the
float32 CPU oracle of the reference's head arithmetic (oracle/heads_oracle.py, pinned by the
fixtures generated from the reference's own classes).

Tolerance: north_star's bf16 bound, rel 2e-2 (``util.rel_err``: max|a-b| / max|b| per tensor) --
the reference itself runs these heads under fp16 autocast
(/root/reference/src/train_new_multimodal_multitask.py:76, src/train_projector.py:49)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import util
from util import RTOL_BF16, rel_err
import mmac_b200 as agx
from mmac_b200 import synth
from oracle import heads_oracle as ho

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _pair(fv, dropout=0.0, seed=11):
    torch.manual_seed(seed)
    orc = ho.MultiTaskHeadOracle(fv, 128, {'style': 32, 'genre': 18}, 0.0)
    util.fill_params_deterministic(orc)
    m = agx.NewMultiModalMultiTaskHead(128, {'style': 32, 'genre': 18}, dropout, feat_size=fv)
    m.load_state_dict(orc.state_dict())
    return orc, m.to(DEV).train()


def _check_multitask(n, arch, fv, weighted, masks):
    feat, emb_s, emb_g, y_s, y_g = synth.make_head_batch(n, arch=arch, seed=5)
    orc, m = _pair(fv, 0.3 if masks else 0.0)
    w_s = synth.class_weights(y_s, 32) if weighted else None
    w_g = synth.class_weights(y_g, 18) if weighted else None
    cs, cg = torch.cat((feat, emb_s), 1), torch.cat((feat, emb_g), 1)
    if masks:
        gen = torch.Generator().manual_seed(9)
        ms = (torch.rand(n, fv + 128, generator=gen) >= 0.3).float() / 0.7
        mg = (torch.rand(n, fv + 128, generator=gen) >= 0.3).float() / 0.7
        cs, cg = cs * ms, cg * mg
        m.dropout_masks = {'style': ms.to(DEV), 'genre': mg.to(DEV)}
    out_o = [orc.class_style[1](cs), orc.class_genre[1](cg)]
    l_o = ho.multitask_loss(out_o, y_s, y_g, w_s, w_g)
    l_o.backward()
    logits = (torch.empty(n, 32, device=DEV), torch.empty(n, 18, device=DEV))
    assert m.tc_supported(feat.to(DEV), emb_s.to(DEV))
    loss = m.train_step_tc(feat.to(DEV), emb_s.to(DEV), emb_g.to(DEV), y_s.to(DEV), y_g.to(DEV),
                           None if w_s is None else w_s.to(DEV), None if w_g is None else w_g.to(DEV),
                           logits=logits)
    torch.cuda.synchronize()
    assert rel_err(logits[0], out_o[0]) <= RTOL_BF16
    assert rel_err(logits[1], out_o[1]) <= RTOL_BF16
    assert rel_err(loss, l_o) <= RTOL_BF16
    assert rel_err(m.class_style[1].weight.grad, orc.class_style[1].weight.grad) <= RTOL_BF16
    assert rel_err(m.class_genre[1].weight.grad, orc.class_genre[1].weight.grad) <= RTOL_BF16
    assert rel_err(m.class_style[1].bias.grad, orc.class_style[1].bias.grad) <= RTOL_BF16
    assert rel_err(m.class_genre[1].bias.grad, orc.class_genre[1].bias.grad) <= RTOL_BF16
    return m


@pytest.mark.parametrize('n', [32, 128, 300, 4096])
@pytest.mark.parametrize('weighted', [False, True])
def test_multitask_step_vit(n, weighted):
    """B = 32 is the reference's default batch (utils.py:23); 300 has a ragged last tile; 4096 the
    benchmarked batch (32 row tiles x 2 heads = 64 CTAs)."""
    _check_multitask(n, 'vit', 768, weighted, masks=False)


def test_multitask_step_explicit_dropout_masks():
    _check_multitask(257, 'vit', 768, True, masks=True)


def test_multitask_step_resnet_split_columns():
    """Fv = 2048: K = 2176 = 17 chunks of 128 columns > the 7 accumulators of one CTA: three CTAs per
    tile repeat the forward pass and own a column range of the weight gradient each."""
    _check_multitask(200, 'resnet', 2048, True, masks=True)


def test_multitask_step_many_tiles_per_cta():
    """More row tiles than CTAs per head (persistent CTAs accumulate their tiles in TMEM)."""
    _check_multitask(128 * 80 + 7, 'vit', 768, False, masks=False)


def test_accumulate_adds_to_existing_gradients_and_is_reproducible():
    n = 200
    feat, emb_s, emb_g, y_s, y_g = [t.to(DEV) for t in synth.make_head_batch(n, 'vit', seed=2)]
    _, m = _pair(768)
    l1 = m.train_step_tc(feat, emb_s, emb_g, y_s, y_g, accumulate=False).clone()
    g1 = m.class_style[1].weight.grad.clone()
    l2 = m.train_step_tc(feat, emb_s, emb_g, y_s, y_g, accumulate=False).clone()
    assert torch.equal(l1, l2) and torch.equal(g1, m.class_style[1].weight.grad)   # fixed order
    m.train_step_tc(feat, emb_s, emb_g, y_s, y_g, accumulate=True)
    assert rel_err(m.class_style[1].weight.grad, 2 * g1) <= 1e-6


def test_philox_dropout_is_consistent_between_forward_and_backward():
    """x = diag: logits[r] = W[:, r] * x[r, r] * mask[r, r] exposes the forward mask of element
    (r, r); d_weight[:, r] = dlogits[r] * x[r, r] * mask[r, r] the mask the second pass applied to
    the same element.  Both must agree, keep with probability 1 - p and scale by 1 / (1 - p)."""
    K, p = 896, 0.4
    n = K
    x = torch.zeros(n, K)
    x[torch.arange(n), torch.arange(n)] = 1.0
    feat, emb = x[:, :768].contiguous().to(DEV), x[:, 768:].contiguous().to(DEV)
    torch.manual_seed(123)
    m = agx.NewMultiModalSingleTaskHead(128, 32, p, feat_size=768).to(DEV).train()
    with torch.no_grad():
        m.classifier[1].weight.fill_(1.0)
        m.classifier[1].bias.zero_()
    y = torch.zeros(n, dtype=torch.int64, device=DEV)
    logits = torch.empty(n, 32, device=DEV)
    m.train_step_tc(feat, emb, y, logits=logits, accumulate=False)
    torch.cuda.synchronize()
    fwd = logits[:, 0]                                       # mask[r, r] / (1 - p)
    kept = fwd > 0
    assert torch.all((fwd == 0) | ((fwd - 1 / (1 - p)).abs() < 2e-2))
    frac = float(kept.float().mean())
    assert abs(frac - (1 - p)) < 0.06, frac
    # dlogits[r, 1] = softmax[1] / n  (class 1 is never the label) is positive: the gradient column r
    # of class 1 is non-zero exactly where the second pass kept element (r, r)
    g = m.classifier[1].weight.grad[1]
    assert torch.equal(g != 0, kept)
    # another step draws another mask; the same seed state reproduces the first one
    logits2 = torch.empty_like(logits)
    m.train_step_tc(feat, emb, y, logits=logits2, accumulate=False)
    assert not torch.equal(logits2[:, 0] > 0, kept)
    m._seed[1] = 0
    m.train_step_tc(feat, emb, y, logits=logits2, accumulate=False)
    assert torch.equal(logits2, logits)
    # eval mode: no dropout
    m.eval()
    m.train_step_tc(feat, emb, y, logits=logits2, accumulate=False)
    assert torch.all((logits2[:, 0] - 1.0).abs() < 1e-2)


@pytest.mark.parametrize('n,fv', [(32, 768), (200, 768), (4096, 768), (130, 2048)])
def test_projector_step(n, fv):
    feat, emb_s, _, _, _ = synth.make_head_batch(n, arch='vit' if fv == 768 else 'resnet', seed=4)
    target = emb_s * 3.0                                     # both SmoothL1 branches
    orc = ho.ProjectorOracle(fv, 128)
    util.fill_params_deterministic(orc)
    m = agx.LabelProjectorHead(128, feat_size=fv)
    m.load_state_dict(orc.state_dict())
    m = m.to(DEV)
    out_o = orc(feat)
    l_o = ho.projector_loss(out_o, target)
    l_o.backward()
    out = torch.empty(n, 128, device=DEV)
    assert m.tc_supported(feat.to(DEV))
    loss = m.train_step_tc(feat.to(DEV), target.to(DEV), out=out, accumulate=False)
    torch.cuda.synchronize()
    assert rel_err(out, out_o) <= RTOL_BF16
    assert rel_err(loss, l_o) <= RTOL_BF16
    assert rel_err(m.encoder.weight.grad, orc.encoder.weight.grad) <= RTOL_BF16
    assert rel_err(m.encoder.bias.grad, orc.encoder.bias.grad) <= RTOL_BF16


def test_heads_golden_bf16():
    """The fixtures generated from the reference's own models_kg.py classes, at the bf16 bound."""
    gold = util.load_golden('heads_vit.npz')
    feat, emb_s, emb_g, y_s, y_g = synth.make_head_batch(96, arch='vit', seed=7)
    m = agx.NewMultiModalMultiTaskHead(128, {'style': 32, 'genre': 18}, 0.0, feat_size=768)
    util.fill_params_deterministic(m)
    m = m.to(DEV).train()
    w_s = synth.class_weights(y_s, 32).to(DEV)
    w_g = synth.class_weights(y_g, 18).to(DEV)
    logits = (torch.empty(96, 32, device=DEV), torch.empty(96, 18, device=DEV))
    loss = m.train_step_tc(feat.to(DEV), emb_s.to(DEV), emb_g.to(DEV), y_s.to(DEV), y_g.to(DEV),
                           w_s, w_g, logits=logits, accumulate=False)
    assert rel_err(logits[0], gold['out_style']) <= RTOL_BF16
    assert rel_err(logits[1], gold['out_genre']) <= RTOL_BF16
    assert rel_err(loss, gold['loss_weighted']) <= RTOL_BF16
    assert rel_err(m.class_style[1].weight.grad, gold['grad_w_style']) <= RTOL_BF16
    assert rel_err(m.class_genre[1].bias.grad, gold['grad_b_genre']) <= RTOL_BF16


def test_head_trainer_bf16_tracks_fp32_trainer():
    """HeadTrainer(precision='bf16') under a CUDA graph: the loss trajectory follows the exact
    float32 trainer (same weights, no dropout) within the bf16 bound."""
    from mmac_b200.trainer import HeadTrainer
    feat, es, eg, ys, yg = [t.to(DEV) for t in synth.make_head_batch(512, 'vit', seed=3)]
    ws = synth.class_weights(ys.cpu(), 32).to(DEV)
    wg = synth.class_weights(yg.cpu(), 18).to(DEV)
    torch.manual_seed(1)
    h32 = agx.NewMultiModalMultiTaskHead(128, {'style': 32, 'genre': 18}, 0.0, 768).to(DEV)
    h16 = agx.NewMultiModalMultiTaskHead(128, {'style': 32, 'genre': 18}, 0.0, 768).to(DEV)
    h16.load_state_dict(h32.state_dict())
    t32 = HeadTrainer(h32, 'multitask', 3e-4, ws, wg, use_cuda_graph=True)
    t16 = HeadTrainer(h16, 'multitask', 3e-4, ws, wg, use_cuda_graph=True, precision='bf16')
    l32 = [float(t32.step(feat, es, eg, ys, yg).item()) for _ in range(6)]
    l16 = [float(t16.step(feat, es, eg, ys, yg).item()) for _ in range(6)]
    assert np.allclose(l16, l32, rtol=RTOL_BF16)
    assert l16[-1] < l16[0]
    p32 = agx.LabelProjectorHead(128, 768).to(DEV)
    p16 = agx.LabelProjectorHead(128, 768).to(DEV)
    p16.load_state_dict(p32.state_dict())
    u32 = HeadTrainer(p32, 'projector', 3e-4, use_cuda_graph=True)
    u16 = HeadTrainer(p16, 'projector', 3e-4, use_cuda_graph=True, precision='bf16')
    a = [float(u32.step(feat, es).item()) for _ in range(6)]
    b = [float(u16.step(feat, es).item()) for _ in range(6)]
    assert np.allclose(b, a, rtol=RTOL_BF16) and b[-1] < b[0]

"""Fusion / projector heads: the arithmetic after the backbone of the reference's
``NewMultiModalMultiTask[ViT]``, ``NewMultiModalSingleTask[Vit]`` and ``LabelProjector[Vit]``
(/root/reference/src/models/models_kg.py:139-280), with the same attribute names and state-dict
keys (``class_style.1.weight``, ``class_genre.1.bias``, ``classifier.1.*``, ``encoder.*``) so a
checkpoint of the reference loads into them and they can replace the ``cat -> Sequential`` lines
(:237-243) inside the reference classes.

``cat(feat, emb) -> Dropout -> Linear`` never materialises the concatenation or the dropped-out
copy: ``feat`` and ``emb`` are two K-segments of one agx GEMM against column slices of the weight,
with the Philox dropout mask applied as the operand is staged into shared memory.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import functional as AF
from . import ops


class _HeadBase(nn.Module):
    def __init__(self):
        super().__init__()
        self._seed = None
        self.dropout_masks: Optional[Dict[str, torch.Tensor]] = None    # test hook

    def _masks(self, name: str, seq: nn.Sequential, parts):
        """One mask per input part (same shape), or None when dropout is inactive."""
        if self.dropout_masks is not None:
            full = self.dropout_masks.get(name)
            if full is None:
                return None
            out, off = [], 0
            for p in parts:
                out.append(full[:, off:off + p.shape[1]].contiguous())
                off += p.shape[1]
            return out
        p_drop = seq[0].p
        if not self.training or p_drop == 0.0:
            return None
        device = parts[0].device
        if self._seed is None or self._seed.device != device:
            s = torch.initial_seed()
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                # batch-sharded heads: every rank draws its own masks (as HeteroModule._seed_state)
                s += 0x9E3779B97F4A7C15 * (dist.get_rank() + 1)
            s &= 0x7fffffffffffffff
            self._seed = torch.tensor([s, 0], dtype=torch.int64, device=device)
        out = []
        for p in parts:
            out.append(ops.dropout_mask(tuple(p.shape), p_drop, self._seed))
            self._seed[1] += (p.numel() + 3) // 4
        return out

    def _head(self, name: str, seq: nn.Sequential, feat, emb):
        lin = seq[1]
        parts = [feat.contiguous(), emb.contiguous()]
        return AF.fused_linear(parts, lin.weight, lin.bias, self._masks(name, seq, parts))

    # ---- fused training step on the tensor cores (agx_head_step) ---------------------------------
    def _tc_seed(self, device):
        if self._seed is None or self._seed.device != device:
            s = torch.initial_seed()
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                s += 0x9E3779B97F4A7C15 * (dist.get_rank() + 1)
            self._seed = torch.tensor([s & 0x7fffffffffffffff, 0], dtype=torch.int64, device=device)
        return self._seed

    def _tc_head_arg(self, name, seq, feat, emb, labels, class_w, coef, logits):
        lin = seq[1]
        mask = None
        if self.dropout_masks is not None:
            mask = self.dropout_masks.get(name)
        return ops.HeadArg(parts=[feat, emb], weight=lin.weight, bias=lin.bias,
                           d_weight=_grad_of(lin.weight), d_bias=_grad_of(lin.bias), loss='ce',
                           labels=labels, class_w=class_w, coef=coef, mask=mask, logits=logits)

    def _tc_run(self, args, seq, group, accumulate):
        """Launch the fused step; dropout = the module's p in training mode (Philox stream of this
        module, one step counter per call), or the injected masks."""
        if getattr(self, '_tc_step', None) is None:
            self._tc_step = ops.HeadStep()
        p = float(seq[0].p) if (self.training and self.dropout_masks is None) else 0.0
        seed = self._tc_seed(args[0].weight.device) if p > 0 else None
        loss = self._tc_step.run(args, p, seed, group=group, accumulate=accumulate)
        if seed is not None:
            seed[1] += 1            # device-side counter: graph-capturable
        return loss.reshape(())


def _grad_of(p: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """The parameter's gradient buffer (FlatAdam's arena view), created zeroed when absent."""
    if p is None:
        return None
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad


class NewMultiModalMultiTaskHead(_HeadBase):
    """models_kg.py:164-193 (ResNet, feat_size=2048) / :217-243 (ViT, feat_size=768)."""

    def __init__(self, emb_size: int, num_classes: Dict[str, int], dropout: float,
                 feat_size: int = 768):
        super().__init__()
        self.class_style = nn.Sequential(nn.Dropout(dropout),
                                         nn.Linear(feat_size + emb_size, num_classes['style']))
        self.class_genre = nn.Sequential(nn.Dropout(dropout),
                                         nn.Linear(feat_size + emb_size, num_classes['genre']))

    def forward(self, visual_features, embedding_style, embedding_genre) -> List[torch.Tensor]:
        out_style = self._head('style', self.class_style, visual_features, embedding_style)
        out_genre = self._head('genre', self.class_genre, visual_features, embedding_genre)
        return [out_style, out_genre]

    def tc_supported(self, visual_features, embedding_style) -> bool:
        return ops.head_step_supported([visual_features, embedding_style], 1) and \
            max(self.class_style[1].out_features, self.class_genre[1].out_features) <= 64

    def train_step_tc(self, visual_features, embedding_style, embedding_genre, style_labels,
                      genre_labels, w_style=None, w_genre=None, group=None, accumulate=True,
                      logits=None) -> torch.Tensor:
        """forward + ``0.5*CE_style + 0.5*CE_genre`` + backward of one mini-batch
        (models_kg.py:237-243, train_new_multimodal_multitask.py:76-83) as ONE tensor-core kernel in
        bf16 with float32 accumulation -- the precision class of the reference's fp16 autocast.
        Adds the gradients to the parameters' ``.grad`` (``accumulate``) and returns the loss (this
        rank's term of the global loss when ``group`` is given).  ``logits``: optional pair of
        ``[B, C]`` float32 output buffers."""
        lg = logits or (None, None)
        args = [self._tc_head_arg('style', self.class_style, visual_features, embedding_style,
                                  style_labels, w_style, 0.5, lg[0]),
                self._tc_head_arg('genre', self.class_genre, visual_features, embedding_genre,
                                  genre_labels, w_genre, 0.5, lg[1])]
        return self._tc_run(args, self.class_style, group, accumulate)


class NewMultiModalSingleTaskHead(_HeadBase):
    """models_kg.py:139-162 / :195-215."""

    def __init__(self, emb_size: int, num_class: int, dropout: float, feat_size: int = 768):
        super().__init__()
        self.classifier = nn.Sequential(nn.Dropout(dropout),
                                        nn.Linear(feat_size + emb_size, num_class))

    def forward(self, visual_features, embedding):
        return self._head('classifier', self.classifier, visual_features, embedding)

    def tc_supported(self, visual_features, embedding) -> bool:
        return ops.head_step_supported([visual_features, embedding], 1) and \
            self.classifier[1].out_features <= 64

    def train_step_tc(self, visual_features, embedding, labels, weight=None, group=None,
                      accumulate=True, logits=None) -> torch.Tensor:
        """Single-task variant (models_kg.py:158-162, train_new_multimodal.py:39-44): one CE."""
        args = [self._tc_head_arg('classifier', self.classifier, visual_features, embedding, labels,
                                  weight, 1.0, logits)]
        return self._tc_run(args, self.classifier, group, accumulate)


class LabelProjectorHead(nn.Module):
    """models_kg.py:245-280: ``encoder = Linear(feat_size, emb_size)``."""

    def __init__(self, emb_size: int, feat_size: int = 768):
        super().__init__()
        self.encoder = nn.Linear(feat_size, emb_size)

    def forward(self, visual_features):
        return AF.fused_linear([visual_features], self.encoder.weight, self.encoder.bias)

    def tc_supported(self, visual_features) -> bool:
        return ops.head_step_supported([visual_features], 1) and \
            self.encoder.out_features <= 64 * ops.L.MAX_HEADS

    def train_step_tc(self, visual_features, embedding, global_rows=None, accumulate=True,
                      out=None) -> torch.Tensor:
        """forward + ``SmoothL1Loss()(encoder(feat), embedding)`` + backward of one mini-batch
        (models_kg.py:261,278, train_projector.py:49-54) on the tensor cores; the 128 outputs run as
        column slices of <= 64.  ``global_rows``: batch rows over all ranks (this rank's term of the
        global mean)."""
        if getattr(self, '_tc_step', None) is None:
            self._tc_step = ops.HeadStep()
        w, b = self.encoder.weight, self.encoder.bias
        dw, db = _grad_of(w), _grad_of(b)
        E = w.shape[0]
        rows = visual_features.shape[0] if global_rows is None else int(global_rows)
        args = []
        for c0 in range(0, E, 64):
            c1 = min(E, c0 + 64)
            args.append(ops.HeadArg(parts=[visual_features], weight=w[c0:c1],
                                    bias=None if b is None else b[c0:c1], d_weight=dw[c0:c1],
                                    d_bias=None if db is None else db[c0:c1], loss='smooth_l1',
                                    target=embedding[:, c0:c1], inv_count=1.0 / (rows * E),
                                    logits=None if out is None else out[:, c0:c1]))
        return self._tc_step.run(args, 0.0, None, accumulate=accumulate).reshape(())


class ContextNetSingleTaskHead(nn.Module):
    """Garcia et al. ContextNet after the backbone (models_kg.py:7-33): ``classifier`` and
    ``encoder`` both read the visual features; returns ``(out, graph_proj)``."""

    def __init__(self, emb_size: int, num_class: int, feat_size: int = 2048):
        super().__init__()
        self.classifier = nn.Linear(feat_size, num_class)
        self.encoder = nn.Linear(feat_size, emb_size)

    def forward(self, visual_features):
        out = AF.fused_linear([visual_features], self.classifier.weight, self.classifier.bias)
        proj = AF.fused_linear([visual_features], self.encoder.weight, self.encoder.bias)
        return out, proj


class ContextNetMultiTaskHead(nn.Module):
    """models_kg.py:35-62: ``class_style`` / ``class_genre`` / ``encoder`` on the visual features;
    returns ``([out_style, out_genre], graph_proj)``."""

    def __init__(self, emb_size: int, num_classes: Dict[str, int], feat_size: int = 2048):
        super().__init__()
        self.class_style = nn.Linear(feat_size, num_classes['style'])
        self.class_genre = nn.Linear(feat_size, num_classes['genre'])
        self.encoder = nn.Linear(feat_size, emb_size)

    def forward(self, visual_features):
        f = [visual_features]
        proj = AF.fused_linear(f, self.encoder.weight, self.encoder.bias)
        return [AF.fused_linear(f, self.class_style.weight, self.class_style.bias),
                AF.fused_linear(f, self.class_genre.weight, self.class_genre.bias)], proj


class _CastellanoBase(_HeadBase):
    """Castellano et al. (models_kg.py:64-137): ``encoder = Linear -> Tanh -> Linear -> Tanh``
    produces the graph projection, the classifiers read ``cat(features, projection)`` through
    ``Dropout(0.2)`` -- concat-free like the new-multimodal heads."""

    def _encode(self, feat):
        e = self.encoder
        h = AF.tanh(AF.fused_linear([feat], e[0].weight, e[0].bias))
        return AF.tanh(AF.fused_linear([h], e[2].weight, e[2].bias))

    @staticmethod
    def _make_encoder(feat_size, emb_size):
        return nn.Sequential(nn.Linear(feat_size, emb_size), nn.Tanh(),
                             nn.Linear(emb_size, emb_size), nn.Tanh())


class MultiModalSingleTaskHead(_CastellanoBase):
    def __init__(self, emb_size: int, num_class: int, feat_size: int = 2048, dropout: float = 0.2):
        super().__init__()
        self.classifier = nn.Sequential(nn.Dropout(dropout),
                                        nn.Linear(feat_size + emb_size, num_class))
        self.encoder = self._make_encoder(feat_size, emb_size)

    def forward(self, visual_features):
        proj = self._encode(visual_features)
        return self._head('classifier', self.classifier, visual_features, proj), proj


class MultiModalMultiTaskHead(_CastellanoBase):
    def __init__(self, emb_size: int, num_classes: Dict[str, int], feat_size: int = 2048,
                 dropout: float = 0.2):
        super().__init__()
        self.class_style = nn.Sequential(nn.Dropout(dropout),
                                         nn.Linear(feat_size + emb_size, num_classes['style']))
        self.class_genre = nn.Sequential(nn.Dropout(dropout),
                                         nn.Linear(feat_size + emb_size, num_classes['genre']))
        self.encoder = self._make_encoder(feat_size, emb_size)

    def forward(self, visual_features):
        proj = self._encode(visual_features)
        return [self._head('style', self.class_style, visual_features, proj),
                self._head('genre', self.class_genre, visual_features, proj)], proj


def context_loss(out, graph_proj, labels, embedding, lamb: float, encoder: str = 'smooth_l1',
                 weight=None, w_genre=None):
    """``lamb * class_loss + (1 - lamb) * encoder_loss`` (src/train_baseline_context.py:47-54,
    75-77; multitask: src/train_baseline_context_multitask.py:76-79 with
    ``class_loss = 0.5*CE_style + 0.5*CE_genre``).  ContextNet: SmoothL1, lamb 0.9 (SGD);
    Castellano: MSE, lamb 0.6 (Adam).  ``out`` / ``labels``: tensors, or [style, genre] pairs."""
    if isinstance(out, (list, tuple)):
        class_loss = AF.cross_entropy(out[0], labels[0], weight, coef=0.5 * lamb) + \
            AF.cross_entropy(out[1], labels[1], w_genre, coef=0.5 * lamb)
    else:
        class_loss = AF.cross_entropy(out, labels, weight, coef=lamb)
    enc = AF.smooth_l1_loss(graph_proj, embedding) if encoder == 'smooth_l1' else \
        AF.mse_loss(graph_proj, embedding)
    return class_loss + (1.0 - lamb) * enc


@torch.no_grad()
def generate_projections(projector: nn.Module, features: torch.Tensor, batch_size: int = 4096):
    """src/generate_projections.py:27-84 after the backbone: run the (trained) projector over all
    validation / test features in batches and return the ``[N, emb]`` tensor the fusion heads index
    (``data_kg.py:169-178``) -- on the device, no file round trip."""
    projector.eval()
    out = torch.empty(features.shape[0], projector.encoder.out_features, dtype=torch.float32,
                      device=projector.encoder.weight.device)
    for i in range(0, features.shape[0], batch_size):
        f = features[i:i + batch_size].to(out.device, non_blocking=True)
        out[i:i + batch_size] = projector(f)
    return out


def multitask_loss(out, style_labels, genre_labels, w_style=None, w_genre=None, group=None):
    """``0.5*CE(out[0], y_style; w) + 0.5*CE(out[1], y_genre; w)``
    (src/train_new_multimodal_multitask.py:48-55,79-81), fused softmax+nll per head.  With
    ``group`` (batch-sharded heads) both weighted means run over the batch shards of all ranks."""
    style_loss = AF.cross_entropy(out[0], style_labels, w_style, coef=0.5, group=group)
    genre_loss = AF.cross_entropy(out[1], genre_labels, w_genre, coef=0.5, group=group)
    return style_loss + genre_loss


def projector_loss(out, embedding, global_rows: Optional[int] = None):
    """``SmoothL1Loss()(out, embedding)`` (src/train_projector.py:33,52).  ``global_rows``: batch
    rows over all ranks -- this rank's term of the global mean (the terms sum to the loss)."""
    loss = AF.smooth_l1_loss(out, embedding)
    if global_rows is not None and global_rows != out.shape[0]:
        loss = loss * (out.shape[0] / float(global_rows))
    return loss


def select_embeddings(table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """a-11 head input selection (src/data/data_kg.py:169-178): ``embedding[idx]`` /
    ``embedding[label_id]`` as one device row gather."""
    return ops.gather_rows(table.contiguous(), idx.to(table.device, torch.int64).contiguous())

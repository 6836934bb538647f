"""CPU restatement of the projector / new-multimodal fusion heads and their losses.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned against the reference's own classes by
tests/golden/make_golden.py (fixtures heads_*.npz).

Follows /root/reference/src/models/models_kg.py:139-280 (head arithmetic after the backbone),
src/train_projector.py:33,49-54 (SmoothL1), src/train_new_multimodal_multitask.py:48-55,76-83
(0.5*CE + 0.5*CE, optional class weights) and src/train_new_multimodal.py:39-44 (single task).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


class MultiTaskHeadOracle(nn.Module):
    """``cat(feat, emb_t) -> Dropout -> Linear`` for t in {style, genre}
    (models_kg.py:174-180,187-193 ResNet / :225-243 ViT).  State-dict keys ``class_style.1.*``,
    ``class_genre.1.*`` as in the reference."""

    def __init__(self, feat_size: int, emb_size: int, num_classes: Dict[str, int], dropout: float):
        super().__init__()
        self.class_style = nn.Sequential(nn.Dropout(dropout),
                                         nn.Linear(feat_size + emb_size, num_classes['style']))
        self.class_genre = nn.Sequential(nn.Dropout(dropout),
                                         nn.Linear(feat_size + emb_size, num_classes['genre']))

    def forward(self, feat, emb_style, emb_genre):
        comb_style = torch.cat((feat, emb_style), dim=1)
        comb_genre = torch.cat((feat, emb_genre), dim=1)
        return [self.class_style(comb_style), self.class_genre(comb_genre)]


class SingleTaskHeadOracle(nn.Module):
    """models_kg.py:148-150,158-162 / :202-215; key ``classifier.1.*``."""

    def __init__(self, feat_size: int, emb_size: int, num_class: int, dropout: float):
        super().__init__()
        self.classifier = nn.Sequential(nn.Dropout(dropout),
                                        nn.Linear(feat_size + emb_size, num_class))

    def forward(self, feat, emb):
        return self.classifier(torch.cat((feat, emb), dim=1))


class ProjectorOracle(nn.Module):
    """models_kg.py:254,261 / :272,278; key ``encoder.*``."""

    def __init__(self, feat_size: int, emb_size: int):
        super().__init__()
        self.encoder = nn.Linear(feat_size, emb_size)

    def forward(self, feat):
        return self.encoder(feat)


def multitask_loss(out, y_style, y_genre, w_style: Optional[torch.Tensor] = None,
                   w_genre: Optional[torch.Tensor] = None):
    """train_new_multimodal_multitask.py:79-81."""
    style_loss = 0.5 * F.cross_entropy(out[0], y_style, weight=w_style)
    genre_loss = 0.5 * F.cross_entropy(out[1], y_genre, weight=w_genre)
    return style_loss + genre_loss


def projector_loss(out, emb):
    """train_projector.py:33,52 -- SmoothL1Loss(), beta=1, mean over B*emb."""
    return F.smooth_l1_loss(out, emb)


def class_weights(labels: torch.Tensor, num_classes: int) -> torch.Tensor:
    """utils.py:268-274: n_total / (n_c * num_classes)."""
    cnt = torch.bincount(labels, minlength=num_classes).to(torch.float64)
    return (cnt.sum() / (cnt * num_classes)).to(torch.float32)


# ---------------------------------------------------------------------------------------------
# ContextNet (Garcia et al.) and Castellano et al. heads after the backbone
# ---------------------------------------------------------------------------------------------
class ContextNetSingleOracle(nn.Module):
    """models_kg.py:7-33: ``classifier`` and ``encoder`` on the pooled ResNet features."""

    def __init__(self, feat_size: int, emb_size: int, num_class: int):
        super().__init__()
        self.classifier = nn.Linear(feat_size, num_class)
        self.encoder = nn.Linear(feat_size, emb_size)

    def forward(self, feat):
        return self.classifier(feat), self.encoder(feat)


class ContextNetMultiOracle(nn.Module):
    """models_kg.py:35-62."""

    def __init__(self, feat_size: int, emb_size: int, num_classes: Dict[str, int]):
        super().__init__()
        self.class_style = nn.Linear(feat_size, num_classes['style'])
        self.class_genre = nn.Linear(feat_size, num_classes['genre'])
        self.encoder = nn.Linear(feat_size, emb_size)

    def forward(self, feat):
        proj = self.encoder(feat)
        return [self.class_style(feat), self.class_genre(feat)], proj


def _castellano_encoder(feat_size, emb_size):
    return nn.Sequential(nn.Linear(feat_size, emb_size), nn.Tanh(), nn.Linear(emb_size, emb_size),
                         nn.Tanh())


class CastellanoSingleOracle(nn.Module):
    """models_kg.py:64-99."""

    def __init__(self, feat_size: int, emb_size: int, num_class: int, dropout: float = 0.2):
        super().__init__()
        self.classifier = nn.Sequential(nn.Dropout(dropout), nn.Linear(feat_size + emb_size, num_class))
        self.encoder = _castellano_encoder(feat_size, emb_size)

    def forward(self, feat):
        proj = self.encoder(feat)
        return self.classifier(torch.cat((feat, proj), 1)), proj


class CastellanoMultiOracle(nn.Module):
    """models_kg.py:101-137."""

    def __init__(self, feat_size: int, emb_size: int, num_classes: Dict[str, int],
                 dropout: float = 0.2):
        super().__init__()
        self.class_style = nn.Sequential(nn.Dropout(dropout),
                                         nn.Linear(feat_size + emb_size, num_classes['style']))
        self.class_genre = nn.Sequential(nn.Dropout(dropout),
                                         nn.Linear(feat_size + emb_size, num_classes['genre']))
        self.encoder = _castellano_encoder(feat_size, emb_size)

    def forward(self, feat):
        proj = self.encoder(feat)
        cat = torch.cat((feat, proj), 1)
        return [self.class_style(cat), self.class_genre(cat)], proj


def context_loss(out, graph_proj, labels, embedding, lamb: float, encoder: str = 'smooth_l1',
                 weight=None, w_genre=None):
    """train_baseline_context.py:47-54,75-77 / train_baseline_context_multitask.py:76-79."""
    if isinstance(out, (list, tuple)):
        class_loss = 0.5 * F.cross_entropy(out[0], labels[0], weight=weight) + \
            0.5 * F.cross_entropy(out[1], labels[1], weight=w_genre)
    else:
        class_loss = F.cross_entropy(out, labels, weight=weight)
    enc = F.smooth_l1_loss(graph_proj, embedding) if encoder == 'smooth_l1' else \
        F.mse_loss(graph_proj, embedding)
    return lamb * class_loss + (1 - lamb) * enc

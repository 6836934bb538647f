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

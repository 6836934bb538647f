"""Deterministic synthetic ArtGraph-shaped heterographs and head feature batches.

The real ArtGraph CSVs are DVC pointers (absent), so every test and benchmark runs on graphs
generated here.  The schema follows the reference's ``ArtGraph.process``
(/root/reference/src/data/artgraph.py:63-117): node-type order, ``eye(N_t)`` one-hot features for
the eight non-artwork types (:93-95), 128-d float32 artwork features (:66-68), float32 label
vectors (:75-81), the nine edge types in the order of :97-105 with ``'_rel'`` appended (:111) and
``edge_index`` as int64 ``[2, E]`` with row 0 = head/source, row 1 = tail/destination (:108-112).

Sizes are those of SURVEY.md section 8(d).  The style / genre class histograms are the row sums of
the reference's published test-set confusion matrices
(results/with_class_weights/new_multimodal_multitask_vit/confusion_matrix_{style,genre}.csv).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

import torch

NODE_TYPES = ['artwork', 'artist', 'gallery', 'style', 'genre', 'tag', 'media', 'field', 'movement']

# (src, rel, dst) in the order of artgraph.py:97-105
EDGE_TYPES = [
    ('artist', 'field_rel', 'field'),
    ('artist', 'movement_rel', 'movement'),
    ('artist', 'teacher_rel', 'artist'),
    ('artwork', 'media_rel', 'media'),
    ('artwork', 'about_rel', 'tag'),
    ('artwork', 'genre_rel', 'genre'),
    ('artwork', 'style_rel', 'style'),
    ('artwork', 'author_rel', 'artist'),
    ('artwork', 'locatedin_rel', 'gallery'),
]

STYLE_HIST = [116, 833, 840, 1460, 168, 171, 206, 286, 440, 219, 187, 638, 239, 224, 437, 118,
              113, 550, 490, 223, 266, 1083, 1595, 211, 2307, 2020, 168, 211, 264, 485, 804, 99]
GENRE_HIST = [434, 351, 277, 278, 2429, 471, 1567, 492, 2754, 3110, 274, 412, 1290, 342, 814,
              517, 708, 951]

SIZES = {
    # cfg_id, node counts, small-relation edge counts
    'tiny': dict(cfg_id=0, artwork=300, artist=24, gallery=9, style=32, genre=18, tag=40,
                 media=7, field=4, movement=6, field_e=14, movement_e=30, teacher_e=12),
    'small': dict(cfg_id=1, artwork=2000, artist=100, gallery=40, style=32, genre=18, tag=200,
                  media=30, field=10, movement=20, field_e=54, movement_e=120, teacher_e=44),
    'mid': dict(cfg_id=3, artwork=14559, artist=313, gallery=137, style=32, genre=18, tag=678,
                media=21, field=7, movement=30, field_e=169, movement_e=375, teacher_e=138),
    'full': dict(cfg_id=2, artwork=116475, artist=2501, gallery=1099, style=32, genre=18,
                 tag=5424, media=167, field=54, movement=243, field_e=1350, movement_e=3000,
                 teacher_e=1100),
}


from .data import EdgeStore, HeteroData, NodeStore  # noqa: E402,F401

HeteroGraph = HeteroData        # the name this module used before the container moved to data.py


def _zipf_probs(n: int, s: float) -> torch.Tensor:
    p = 1.0 / torch.arange(1, n + 1, dtype=torch.float64) ** s
    return p / p.sum()


def _draw(p: torch.Tensor, n: int, gen: torch.Generator) -> torch.Tensor:
    if n == 0:
        return torch.zeros(0, dtype=torch.int64)
    return torch.multinomial(p, n, replacement=True, generator=gen)


def _shuffle(ei: torch.Tensor, gen: torch.Generator) -> torch.Tensor:
    perm = torch.randperm(ei.shape[1], generator=gen)
    return ei[:, perm].contiguous()


def make_artgraph(size: str = 'small', features: str = 'one-hot', seed: int | None = None,
                  feat_dim: int = 128) -> HeteroGraph:
    """Build one synthetic ArtGraph-shaped directed heterograph (before ``ToUndirected``).

    ``features='one-hot'`` is the reference's layout (``eye(N_t)`` for non-artwork types);
    ``features='dense'`` gives every type ``randn(N_t, feat_dim)`` (isolates the kernels from the
    one-hot artefact).
    """
    cfg = SIZES[size]
    gen = torch.Generator().manual_seed(1234 + cfg['cfg_id'] if seed is None else seed)
    A = cfg['artwork']
    g = HeteroGraph()
    g['artwork'].x = torch.randn(A, feat_dim, generator=gen, dtype=torch.float32)

    style_p = torch.tensor(STYLE_HIST, dtype=torch.float64)
    genre_p = torch.tensor(GENRE_HIST, dtype=torch.float64)
    style = _draw(style_p / style_p.sum(), A, gen)
    genre = _draw(genre_p / genre_p.sum(), A, gen)
    g['artwork'].y_style = style.to(torch.float32)
    g['artwork'].y_genre = genre.to(torch.float32)

    for t in NODE_TYPES[1:]:
        n = cfg[t]
        if features == 'one-hot':
            g[t].x = torch.eye(n, dtype=torch.float32)
        elif features == 'identity':        # the same features as a marker (data.Identity)
            from .data import Identity
            g[t].x = Identity(n)
        elif features == 'dense':
            g[t].x = torch.randn(n, feat_dim, generator=gen, dtype=torch.float32)
        else:
            raise ValueError(features)

    aw = torch.arange(A, dtype=torch.int64)
    n_artist = cfg['artist']

    def rand_artists(n):
        return torch.randint(0, n_artist, (n,), generator=gen, dtype=torch.int64)

    edges = {}
    edges['field_rel'] = torch.stack([rand_artists(cfg['field_e']),
                                      _draw(_zipf_probs(cfg['field'], 1.0), cfg['field_e'], gen)])
    edges['movement_rel'] = torch.stack([rand_artists(cfg['movement_e']),
                                         _draw(_zipf_probs(cfg['movement'], 1.0),
                                               cfg['movement_e'], gen)])
    # teacher edges may repeat and may be self-loops: exercises the coalesce path of ToUndirected
    edges['teacher_rel'] = torch.stack([rand_artists(cfg['teacher_e']),
                                        rand_artists(cfg['teacher_e'])])

    n_media = torch.poisson(torch.full((A,), 0.9, dtype=torch.float64), generator=gen).long()
    edges['media_rel'] = torch.stack([aw.repeat_interleave(n_media),
                                      _draw(_zipf_probs(cfg['media'], 1.0),
                                            int(n_media.sum()), gen)])
    n_tag = torch.poisson(torch.full((A,), 3.35, dtype=torch.float64), generator=gen).long()
    edges['about_rel'] = torch.stack([aw.repeat_interleave(n_tag),
                                      _draw(_zipf_probs(cfg['tag'], 1.0), int(n_tag.sum()), gen)])
    edges['genre_rel'] = torch.stack([aw, genre])
    edges['style_rel'] = torch.stack([aw, style])
    edges['author_rel'] = torch.stack([aw, _draw(_zipf_probs(n_artist, 1.1), A, gen)])
    in_gallery = torch.rand(A, generator=gen, dtype=torch.float64) < 0.386
    ga = aw[in_gallery]
    edges['locatedin_rel'] = torch.stack([ga, _draw(_zipf_probs(cfg['gallery'], 1.0),
                                                    ga.numel(), gen)])

    for (s, r, d) in EDGE_TYPES:
        g[(s, r, d)].edge_index = _shuffle(edges[r].to(torch.int64), gen)
    return g


def write_artgraph_raw(root: str, g: HeteroGraph) -> str:
    """Write ``g`` (a directed graph from ``make_artgraph``) as the raw CSV tree the reference's
    ``ArtGraph.process`` reads (/root/reference/src/data/artgraph.py:63-112), under ``root/raw``:

        node-feat/artwork/node-feat.csv          A x 128 floats, no header           (:66-68)
        node-label/artwork/node-label-{style,genre}.csv   one label per line         (:75-81)
        num-node-dict.csv                        header = node types, one row        (:84-95)
        relations/<h>___<r>___<t>/edge.csv       E x 2 ints (head, tail), no header  (:97-112)

    so that the unmodified dataset class builds the same ``HeteroData`` from disk."""
    import os
    import numpy as np
    raw = os.path.join(root, 'raw')
    for sub in ('node-feat/artwork', 'node-label/artwork', 'relations'):
        os.makedirs(os.path.join(raw, sub), exist_ok=True)
    np.savetxt(os.path.join(raw, 'node-feat', 'artwork', 'node-feat.csv'),
               g['artwork'].x.numpy(), delimiter=',', fmt='%.9g')
    for lab in ('style', 'genre'):
        np.savetxt(os.path.join(raw, 'node-label', 'artwork', f'node-label-{lab}.csv'),
                   g['artwork'][f'y_{lab}'].numpy().astype(np.int64), fmt='%d')
    n = g.num_nodes_dict
    with open(os.path.join(raw, 'num-node-dict.csv'), 'w') as fh:
        fh.write(','.join(NODE_TYPES) + '\n' + ','.join(str(n[t]) for t in NODE_TYPES) + '\n')
    for (h, r, t) in EDGE_TYPES:
        d = os.path.join(raw, 'relations', '___'.join((h, r[:-len('_rel')], t)))
        os.makedirs(d, exist_ok=True)
        np.savetxt(os.path.join(d, 'edge.csv'), g[(h, r, t)].edge_index.t().numpy(),
                   delimiter=',', fmt='%d')
    return raw


def replicate(g: HeteroGraph, copies: int) -> HeteroGraph:
    """Block-diagonal ``copies``-fold replication (config 5): node ids of copy c are offset by
    ``c * N_t``; features are tiled so input widths stay unchanged."""
    out = HeteroGraph()
    n = g.num_nodes_dict
    for t in g.node_types:
        for a, v in g[t].items():
            out[t][a] = v.repeat(copies, *([1] * (v.dim() - 1))) if torch.is_tensor(v) else v
    for (s, r, d) in g.edge_types:
        ei = g[(s, r, d)].edge_index
        parts = []
        for c in range(copies):
            off = torch.tensor([[c * n[s]], [c * n[d]]], dtype=torch.int64, device=ei.device)
            parts.append(ei + off)
        out[(s, r, d)].edge_index = torch.cat(parts, dim=1).contiguous()
    return out


def make_head_batch(n: int, arch: str = 'vit', seed: int = 1, emb_dim: int = 128):
    """Synthetic inputs of the fusion / projector heads: precomputed backbone features
    (768-d ViT CLS / 2048-d ResNet50 pooled), style and genre node embeddings and labels drawn
    from the class histograms."""
    fv = 768 if arch == 'vit' else 2048
    gen = torch.Generator().manual_seed(seed)
    feat = torch.randn(n, fv, generator=gen, dtype=torch.float32)
    emb_s = torch.randn(n, emb_dim, generator=gen, dtype=torch.float32)
    emb_g = torch.randn(n, emb_dim, generator=gen, dtype=torch.float32)
    sp = torch.tensor(STYLE_HIST, dtype=torch.float64)
    gp = torch.tensor(GENRE_HIST, dtype=torch.float64)
    y_s = _draw(sp / sp.sum(), n, gen)
    y_g = _draw(gp / gp.sum(), n, gen)
    return feat, emb_s, emb_g, y_s, y_g


def class_weights(labels: torch.Tensor, num_classes: int) -> torch.Tensor:
    """Inverse-frequency class weights ``n_total / (n_c * num_classes)``
    (/root/reference/src/utils.py:268-274)."""
    cnt = torch.bincount(labels, minlength=num_classes).to(torch.float64)
    return (cnt.sum() / (cnt.clamp(min=1) * num_classes)).to(torch.float32)

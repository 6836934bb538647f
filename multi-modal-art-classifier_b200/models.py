"""``HeteroGNN`` / ``HeteroSGNN``: same constructor signatures, sub-module names (``convs``,
``lins``, ``bns``, ``conv_out``, ``gnn``) and dataflow as the reference's model definitions
(/root/reference/src/models/models_graph.py:5-49), built on the agx operators, so that
state-dicts and call sites (``train_gnn_embeddings.py:128-137,42,57-58,89``) carry over.

Dataflow that is reproduced on purpose:
  * what a block hands to the next block -- and what is finally returned as the node embedding --
    is the block's *pre-activation* output (after BatchNorm);
  * activation followed by dropout is computed per block but only the last block's result is
    consumed (by ``conv_out``);
  * the ``training`` flag is read while ``to_hetero`` traces the module, i.e. once: dropout stays
    active after ``.eval()``, BatchNorm (a module call) does honour ``.eval()``.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.nn import functional as F

from . import nn as agx_nn
from .hetero import to_hetero


class HeteroGNN(nn.Module):
    """Homogeneous template that ``to_hetero`` expands per node / edge type."""

    def __init__(self, operator=agx_nn.SAGEConv, activation=nn.ReLU, hidden_channels=128,
                 out_channels=300, num_layers=1, dropout=0.5, bn=False, skip=False):
        super().__init__()
        self.dropout, self.bn, self.skip = dropout, bn, skip
        self.activation = activation
        self.convs, self.lins, self.bns = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for _layer in range(num_layers):
            self.convs.append(operator((-1, -1), hidden_channels))      # lazy input widths
            self.lins.append(agx_nn.Linear(-1, hidden_channels))        # only used when skip=True
            self.bns.append(nn.BatchNorm1d(hidden_channels))
        self.conv_out = operator((-1, -1), out_channels)

    def _block(self, i: int, h, edge_index):
        out = self.convs[i](h, edge_index)
        if self.skip:
            out = out + self.lins[i](h)
        return self.bns[i](out) if self.bn else out

    def forward(self, x, edge_index):
        h, h_act = x, None
        for i in range(len(self.convs)):
            h = self._block(i, h, edge_index)
            h_act = self.activation(h)
            if self.training:                       # frozen at trace time
                h_act = F.dropout(h_act, self.dropout)
        logits = self.conv_out(h_act, edge_index)
        return h, F.log_softmax(logits, dim=1)


class HeteroSGNN(nn.Module):
    """``forward(x_dict, edge_index_dict) -> (embedding_dict, [log_prob_dict])``."""

    def __init__(self, operator, activation, aggr, hidden_channels, out_channels, metadata,
                 n_layers, dropout, bn, skip):
        super().__init__()
        template = HeteroGNN(operator, activation, hidden_channels, out_channels, n_layers,
                             dropout, bn, skip)
        self.gnn = to_hetero(template, metadata, aggr=aggr)

    def forward(self, x, edge_index):
        emb, out_soft = self.gnn(x, edge_index)
        return emb, [out_soft]


class HeteroMGNN(nn.Module):
    """Three independent hetero towers with per-task output widths (artist / style / genre), as
    /root/reference/src/models/models_graph.py:51-64; ``out_channels`` is a dict with those keys.
    ``forward`` returns ``[tower(x, ei) for tower in (artist, style, genre)]``, each an
    ``(embedding_dict, log_prob_dict)`` pair."""

    def __init__(self, operator, activation, aggr, hidden_channels, out_channels, metadata,
                 n_layers, dropout, bn, skip):
        super().__init__()
        def tower(c):
            return to_hetero(HeteroGNN(operator, activation, hidden_channels, c, n_layers, dropout,
                                       bn, skip), metadata, aggr=aggr)
        self.gnn_artist = tower(out_channels['artist'])
        self.gnn_style = tower(out_channels['style'])
        self.gnn_genre = tower(out_channels['genre'])

    def forward(self, x, edge_index):
        return [self.gnn_artist(x, edge_index), self.gnn_style(x, edge_index),
                self.gnn_genre(x, edge_index)]

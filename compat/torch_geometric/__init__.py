"""Import-compatible stand-in for the part of PyG 2.0.2 that the reference's GNN path imports
(``import torch_geometric.nn as operators``, ``import torch_geometric.transforms as T``,
``from torch_geometric.data import InMemoryDataset, HeteroData`` --
/root/reference/src/models/models_graph.py:3, src/train_gnn_embeddings.py:3-4,
src/data/artgraph.py:5-8), bound to the B200 implementation in ``mmac_b200``.

Put this directory's parent (``compat/``) and the repository root on ``PYTHONPATH`` and the
reference's ``models_graph.py`` / ``train_gnn_embeddings.py`` / ``artgraph.py`` import and run as
they are: operators and ``to_hetero`` execute on the GPU (libagx.so), the host tensors the
CPU-only script passes are staged on the device once, results come back to the host lazily
(``to_hetero(host_io=True)``).  Nothing here is a CPU implementation."""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)

import mmac_b200 as _agx  # noqa: E402,F401

from . import data, nn, transforms  # noqa: E402,F401

__version__ = '2.0.2+agx'

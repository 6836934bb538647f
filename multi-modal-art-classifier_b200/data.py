"""``HeteroData``: the heterogeneous graph container the reference's scripts use
(``torch_geometric.data.HeteroData``; PyG is absent in this image).  Only the surface the path
touches is provided -- /root/reference/src/data/artgraph.py:64-112 (``data['artwork'].x = ...``,
``data[(h, r, t)].edge_index = ...``), src/train_gnn_embeddings.py:42,57-58,68-80,133,139-142
(``x_dict``, ``edge_index_dict``, ``metadata()``, ``del data[...]``, ``data['artwork']['y_style']``)
-- plus ``InMemoryDataset`` with the download / process / collate protocol ``ArtGraph`` relies on
(artgraph.py:31-40,59-117)."""
from __future__ import annotations

import copy
import os
from collections import OrderedDict
from typing import Dict, Tuple

import torch


class _Store:
    """Attribute bag of one node or edge type: ``store.x`` and ``store['x']`` are the same.  A plain
    object (attributes in ``__dict__``), so that ``torch.save`` / ``torch.load(weights_only=True)``
    round-trip it (ArtGraph.process / __init__, artgraph.py:40,117)."""

    def __init__(self, mapping=None):
        if mapping is not None:
            self.__dict__.update(dict(mapping.items()))

    def __getitem__(self, key):
        return self.__dict__[key]

    def __setitem__(self, key, value):
        self.__dict__[key] = value

    def __delitem__(self, key):
        del self.__dict__[key]

    def __contains__(self, key):
        return key in self.__dict__

    def __iter__(self):
        return iter(self.__dict__)

    def __len__(self):
        return len(self.__dict__)

    def keys(self):
        return self.__dict__.keys()

    def values(self):
        return self.__dict__.values()

    def items(self):
        return self.__dict__.items()

    def get(self, key, default=None):
        return self.__dict__.get(key, default)

    def __repr__(self):
        return f'{type(self).__name__}({", ".join(self.__dict__)})'


class NodeStore(_Store):
    pass


class EdgeStore(_Store):
    pass


class HeteroData:
    """Insertion-ordered node and edge stores (the order ``to_hetero`` and ``ToUndirected`` see)."""

    def __init__(self):
        self._nodes: "OrderedDict[str, NodeStore]" = OrderedDict()
        self._edges: "OrderedDict[Tuple[str, str, str], EdgeStore]" = OrderedDict()

    def __getitem__(self, key):
        if isinstance(key, tuple):
            key = tuple(key)
            if key not in self._edges:
                self._edges[key] = EdgeStore()
            return self._edges[key]
        if key not in self._nodes:
            self._nodes[key] = NodeStore()
        return self._nodes[key]

    def __delitem__(self, key):
        if isinstance(key, tuple):
            del self._edges[tuple(key)]
        else:
            del self._nodes[key]

    def __contains__(self, key):
        return key in self._edges if isinstance(key, tuple) else key in self._nodes

    @property
    def node_types(self):
        return list(self._nodes.keys())

    @property
    def edge_types(self):
        return list(self._edges.keys())

    def metadata(self):
        return self.node_types, self.edge_types

    @property
    def x_dict(self) -> Dict[str, torch.Tensor]:
        return OrderedDict((k, v['x']) for k, v in self._nodes.items() if 'x' in v)

    @property
    def edge_index_dict(self):
        return OrderedDict((k, v['edge_index']) for k, v in self._edges.items()
                           if 'edge_index' in v)

    @property
    def num_nodes_dict(self):
        out = OrderedDict()
        for k, v in self._nodes.items():
            out[k] = int(v['x'].shape[0]) if 'x' in v else int(v['num_nodes'])
        return out

    def num_edges(self) -> int:
        return sum(int(v['edge_index'].shape[1]) for v in self._edges.values())

    def to(self, device, non_blocking=False):
        g = type(self)()
        for k, v in self._nodes.items():
            for a, t in v.items():
                g[k][a] = t.to(device, non_blocking=non_blocking) if torch.is_tensor(t) else t
        for k, v in self._edges.items():
            for a, t in v.items():
                g[k][a] = t.to(device, non_blocking=non_blocking) if torch.is_tensor(t) else t
        return g

    def __copy__(self):
        g = type(self)()
        for k, v in self._nodes.items():
            g._nodes[k] = NodeStore(v)
        for k, v in self._edges.items():
            g._edges[k] = EdgeStore(v)
        return g

    def __repr__(self):
        return (f'HeteroData(nodes={dict(self.num_nodes_dict)}, '
                f'edges={{{", ".join(f"{k}: {int(v.edge_index.shape[1])}" for k, v in self._edges.items() if "edge_index" in v)}}})')


class Identity:
    """Marker for one-hot node features: ``x_dict['tag'] = Identity(5424)`` stands for
    ``torch.eye(5424)`` (/root/reference/src/data/artgraph.py:93-95) without the caller having to
    build, hold or upload the N x N matrix.  ``to_hetero`` modules and ``GNNTrainer`` accept it
    wherever a feature tensor is accepted: ``I W^T = W^T`` is used directly and no check of the
    values is needed.  A dense ``torch.eye`` tensor keeps working (it is detected on the device)."""

    def __init__(self, n: int):
        self.n = int(n)

    @property
    def shape(self):
        return (self.n, self.n)

    def __repr__(self):
        return f'Identity({self.n})'


_EYES: dict = {}
_DECLARED: set = set()


def identity_tensor(n: int, device) -> torch.Tensor:
    """The device tensor that stands behind ``Identity(n)`` (one per size and device, written on
    the device -- nothing crosses PCIe); kernels never read it when the features are declared
    one-hot, it only carries the shape through the module."""
    key = (int(n), str(device))
    t = _EYES.get(key)
    if t is None:
        t = torch.eye(int(n), dtype=torch.float32, device=device)
        _EYES[key] = t
        _DECLARED.add(t.data_ptr())
    return t


def is_declared_identity(x: torch.Tensor) -> bool:
    return torch.is_tensor(x) and x.data_ptr() in _DECLARED and x._version == 0


def _as_list(x):
    return [x] if isinstance(x, str) else list(x)


class InMemoryDataset:
    """The protocol of PyG 2.0.x ``InMemoryDataset`` that ``ArtGraph`` uses
    (/root/reference/src/data/artgraph.py:31-40): ``raw_dir`` / ``processed_dir`` under ``root``,
    ``download()`` when a raw file is missing, ``process()`` when a processed file is missing,
    ``collate`` of a one-element list, ``dataset[0]`` with the optional ``transform``."""

    def __init__(self, root=None, transform=None, pre_transform=None, pre_filter=None):
        self.root = os.path.expanduser(os.path.normpath(root)) if isinstance(root, str) else root
        self.transform, self.pre_transform, self.pre_filter = transform, pre_transform, pre_filter
        self.data, self.slices = None, None
        if self.root is not None:
            self._download()
            self._process()

    @property
    def raw_dir(self):
        return os.path.join(self.root, 'raw')

    @property
    def processed_dir(self):
        return os.path.join(self.root, 'processed')

    @property
    def raw_file_names(self):
        raise NotImplementedError

    @property
    def processed_file_names(self):
        raise NotImplementedError

    @property
    def raw_paths(self):
        return [os.path.join(self.raw_dir, f) for f in _as_list(self.raw_file_names)]

    @property
    def processed_paths(self):
        return [os.path.join(self.processed_dir, f) for f in _as_list(self.processed_file_names)]

    def download(self):
        raise NotImplementedError

    def process(self):
        raise NotImplementedError

    def _download(self):
        if all(os.path.exists(f) for f in self.raw_paths):
            return
        os.makedirs(self.raw_dir, exist_ok=True)
        self.download()

    def _process(self):
        if all(os.path.exists(f) for f in self.processed_paths):
            return
        os.makedirs(self.processed_dir, exist_ok=True)
        self.process()

    @staticmethod
    def collate(data_list):
        if len(data_list) != 1:
            raise NotImplementedError('collate of more than one graph (ArtGraph stores one)')
        return data_list[0], None

    def __len__(self):
        return 1

    def __getitem__(self, idx):
        if idx not in (0, -1):
            raise IndexError(idx)
        data = copy.copy(self.data)
        return data if self.transform is None else self.transform(data)


# ``ArtGraph.__init__`` reads its processed file back with a bare ``torch.load`` (artgraph.py:40);
# torch >= 2.6 unpickles with weights_only=True and needs the container classes allow-listed
try:
    torch.serialization.add_safe_globals([HeteroData, NodeStore, EdgeStore, OrderedDict])
except AttributeError:      # older torch: plain pickle
    pass

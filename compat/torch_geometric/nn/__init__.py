"""``torch_geometric.nn`` names used by the reference (models_graph.py:17-18,45;
train_gnn_embeddings.py:96-102)."""
import mmac_b200 as _agx
from mmac_b200.nn import GATConv, GraphConv, Linear, MessagePassing, SAGEConv  # noqa: F401


def to_hetero(module, metadata, aggr='sum', **kwargs):
    """``operators.to_hetero(gnn, metadata, aggr=aggr)`` (models_graph.py:45).  ``host_io``: the
    reference script feeds host tensors and computes its loss on the host
    (train_gnn_embeddings.py:30,42)."""
    kwargs.setdefault('host_io', True)
    return _agx.to_hetero(module, metadata, aggr, **kwargs)


class _NotBipartite:
    """GCNConv / GINConv are REGISTERED by the script (train_gnn_embeddings.py:100-101) but cannot
    be built as ``operator((-1, -1), hidden)`` on ArtGraph's bipartite edge types in PyG 2.0.2
    either: the names exist, constructing one explains that."""
    _name = ''

    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            f'{self._name}: not constructible as operator((-1, -1), channels) on bipartite edge '
            f'types (also in PyG 2.0.2); use SAGEConv, GraphConv or GATConv')


class GCNConv(_NotBipartite):
    _name = 'GCNConv'


class GINConv(_NotBipartite):
    _name = 'GINConv'

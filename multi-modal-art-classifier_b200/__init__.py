"""B200-native (sm_100a) implementation of the hetero-GNN + fusion-head hot path of
CILAB-ArtGraph/multi-modal-art-classifier.  Hand-written CUDA behind a C ABI (include/agx.h,
``libagx.so``); torch supplies device memory, streams and ``torch.distributed`` only.

    import mmac_b200 as agx
    data = agx.ToUndirected()(data)                       # T.ToUndirected()
    model = agx.HeteroSGNN(agx.SAGEConv, torch.nn.ReLU(), 'sum', 128, 32, data.metadata(),
                           n_layers=2, dropout=0.4, bn=True, skip=False)
"""
__version__ = "0.1.0"

from . import synth  # noqa: F401
from ._lib import AgxError, EXPORTED, lib  # noqa: F401
from .data import HeteroData, Identity, InMemoryDataset  # noqa: F401
from .graph import HeteroPlan, ToUndirected, get_plan, to_undirected_dict  # noqa: F401
from .nn import GATConv, GraphConv, Linear, MessagePassing, SAGEConv  # noqa: F401
from .hetero import HeteroModule, to_hetero  # noqa: F401
from .models import HeteroGNN, HeteroMGNN, HeteroSGNN  # noqa: F401
from .heads import (ContextNetMultiTaskHead, ContextNetSingleTaskHead,  # noqa: F401
                    LabelProjectorHead, MultiModalMultiTaskHead, MultiModalSingleTaskHead,
                    NewMultiModalMultiTaskHead, NewMultiModalSingleTaskHead, context_loss,
                    generate_projections, multitask_loss, projector_loss, select_embeddings)
from .optim import FlatAdam, FlatSGD  # noqa: F401
from . import functional  # noqa: F401

"""``torch_geometric.transforms`` names used by the reference (train_gnn_embeddings.py:117-120)."""
from mmac_b200.graph import ToUndirected  # noqa: F401

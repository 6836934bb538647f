"""B200-native hetero-GNN + fusion-head hot path of CILAB-ArtGraph/multi-modal-art-classifier."""
__version__ = "0.1.0"

"""``torch_geometric.data`` names imported by /root/reference/src/data/artgraph.py:7-8."""
from mmac_b200.data import EdgeStore, HeteroData, InMemoryDataset, NodeStore  # noqa: F401


def download_url(*args, **kwargs):
    raise RuntimeError('download_url: no network access; place the raw files under <root>/raw')


def extract_zip(*args, **kwargs):
    raise RuntimeError('extract_zip: not provided (ArtGraph never calls it)')

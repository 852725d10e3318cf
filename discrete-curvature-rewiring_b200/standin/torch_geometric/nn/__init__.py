"""``torch_geometric.nn`` stand-in: only ``GCNConv`` (what ``models/gcn.py:9,19,36`` of the reference uses)."""
from .gcn_conv import GCNConv  # noqa: F401

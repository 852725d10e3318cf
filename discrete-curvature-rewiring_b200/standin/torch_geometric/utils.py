"""The five ``torch_geometric.utils`` functions imported at ``rewiring/sdrf_cuda_bfc.py:7`` (PyG 2.0.3 semantics)."""
from __future__ import annotations

import torch

from .data import Data


def _maybe_num_nodes(edge_index, num_nodes=None):
    if num_nodes is not None:
        return num_nodes
    return int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0


def remove_self_loops(edge_index, edge_attr=None):
    mask = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, mask]
    return (edge_index, None) if edge_attr is None else (edge_index, edge_attr[mask])


def to_undirected(edge_index, edge_attr=None, num_nodes=None, reduce="add"):
    """Concatenate with the flipped copy and coalesce (sort by ``row*N+col``, drop duplicates)."""
    if edge_attr is not None:
        raise NotImplementedError("stand-in: edge_attr is not used by the reference hot path")
    n = _maybe_num_nodes(edge_index, num_nodes)
    row, col = edge_index[0], edge_index[1]
    row, col = torch.cat([row, col]), torch.cat([col, row])
    key = torch.unique(row * n + col, sorted=True)
    return torch.stack([torch.div(key, n, rounding_mode="floor"), key % n])


def to_dense_adj(edge_index, batch=None, edge_attr=None, max_num_nodes=None):
    """``[1, N, N]`` fp32, ``N = edge_index.max()+1``; duplicate edges sum."""
    if batch is not None or edge_attr is not None:
        raise NotImplementedError("stand-in: batch / edge_attr are not used by the reference hot path")
    n = _maybe_num_nodes(edge_index, max_num_nodes)
    adj = torch.zeros(n * n, dtype=torch.float32, device=edge_index.device)
    idx = edge_index[0] * n + edge_index[1]
    adj.index_add_(0, idx, torch.ones(idx.numel(), dtype=torch.float32, device=edge_index.device))
    return adj.view(1, n, n)


def to_networkx(data, node_attrs=None, edge_attrs=None, to_undirected=False, remove_self_loops=False):
    import networkx as nx

    G = nx.Graph() if to_undirected else nx.DiGraph()
    G.add_nodes_from(range(data.num_nodes))
    node_attrs = node_attrs or []
    values = {k: getattr(data, k).squeeze().tolist() if torch.is_tensor(getattr(data, k)) else getattr(data, k)
              for k in node_attrs}
    for u, v in data.edge_index.t().tolist():
        if to_undirected and v > u:
            continue
        if remove_self_loops and u == v:
            continue
        G.add_edge(u, v)
    for key in node_attrs:
        for i, feat in G.nodes(data=True):
            feat.update({key: values[key][i]})
    return G


def from_networkx(G):
    import networkx as nx

    G = nx.convert_node_labels_to_integers(G)
    G = G.to_directed() if not nx.is_directed(G) else G
    edges = list(G.edges)
    edge_index = torch.tensor(edges, dtype=torch.long).t().contiguous().view(2, -1)
    data = Data(edge_index=edge_index)
    node_keys = set()
    for _, feat in G.nodes(data=True):
        node_keys |= set(feat.keys())
    for key in node_keys:
        try:
            setattr(data, str(key), torch.tensor([feat[key] for _, feat in G.nodes(data=True)]))
        except (ValueError, TypeError, KeyError):
            pass
    data.num_nodes = G.number_of_nodes()
    return data

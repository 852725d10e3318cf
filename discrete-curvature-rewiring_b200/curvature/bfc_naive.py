"""Drop-in for the reference's ``curvature/bfc_naive.py``: paper-definition Balanced Forman curvature.

``bfc_edge(G, v1, v2) -> float``      reference: curvature/bfc_naive.py:7-40
``bfc(G) -> G``                       reference: curvature/bfc_naive.py:43-52   (sets ``G[v1][v2]['bfc']``)

The reference walks Python sets per edge and rebuilds a scipy sparse matrix per edge (:34).  Here the graph goes
to the GPU once as a sorted CSR and one launch of the paper-flavour kernel of ``libdcr.so`` produces, for every
edge, #triangles, #squares at each endpoint, gamma_max and the fp64 curvature (evaluated in the reference's
left-to-right order).  Node labels must be the integers ``0..N-1`` — the reference indexes adjacency-matrix rows
by label (:36-37), so it assumes the same.
"""
import networkx as nx
import numpy as np
import torch

from dcr import bfc as _bfc
from dcr import graph as _graph


def _csr_of(G: nx.Graph) -> "_bfc.DeviceCSR":
    n = G.number_of_nodes()
    if n and (min(G.nodes) != 0 or max(G.nodes) != n - 1):
        raise ValueError("node labels must be the integers 0..N-1")
    e = np.array([(u, v) for u, v in G.edges() if u != v], dtype=np.int64).reshape(-1, 2)
    rowptr, col = _graph.undirected_csr(e.T, num_nodes=n)
    return _bfc.DeviceCSR.from_host(rowptr, col)


def _run(csr, src, dst):
    esrc = torch.as_tensor(src, dtype=torch.int32).to(csr.colidx.device)
    edst = torch.as_tensor(dst, dtype=torch.int32).to(csr.colidx.device)
    out = _bfc.paper_flavour(csr, edges=(esrc, edst))
    return out


class PreparedGraph:
    """``prepare(G)``: the device-resident CSR of ``G``, for callers that query many edges of an UNCHANGED graph with
    ``bfc_edge``.  A networkx graph carries no modification counter, so ``bfc_edge(G, …)`` on a plain graph has to rebuild
    the CSR on every call (O(E), like the reference's per-call ``nx.adj_matrix(G)``, bfc_naive.py:34); passing the
    prepared object instead makes a call O(its own kernel)."""

    def __init__(self, G: nx.Graph):
        self.csr = _csr_of(G)
        self.degree = dict(G.degree)


def prepare(G: nx.Graph) -> PreparedGraph:
    return PreparedGraph(G)


def bfc_edge(G, v1: int, v2: int) -> float:
    """Balanced Forman curvature of the edge ``(v1, v2)`` of the undirected graph ``G`` (an ``nx.Graph``, or the result of
    :func:`prepare` for repeated queries on an unchanged graph)."""
    deg = G.degree
    if min(deg[v1], deg[v2]) == 1:                # bfc_naive.py:18-19 returns the int 0
        return 0
    out = _run(G.csr if isinstance(G, PreparedGraph) else _csr_of(G), [v1], [v2])
    return float(out["bfc"].cpu()[0])


def bfc(G: nx.Graph) -> nx.Graph:
    """Assign ``G[v1][v2]['bfc']`` for every edge (one kernel launch for the whole graph)."""
    edges = [(u, v) for u, v in G.edges]
    if not edges:
        return G
    out = _run(_csr_of(G), [u for u, _ in edges], [v for _, v in edges])
    vals = out["bfc"].cpu().tolist()
    for (u, v), val in zip(edges, vals):
        G[u][v]['bfc'] = 0 if min(G.degree[u], G.degree[v]) == 1 else val
    return G

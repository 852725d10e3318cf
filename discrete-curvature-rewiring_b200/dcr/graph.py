"""Host-side graph set-up for the SDRF / BFC entry points (numpy; runs once per call, not in the hot loop).

Restates, vectorised, the third-party canonicalisation the reference relies on:
  * ``to_undirected`` + ``remove_self_loops``           (rewiring/sdrf_cuda_bfc.py:26-29)  -> :func:`undirected_csr`
  * ``to_networkx(data).to_undirected()`` insertion order (:31-33)                          -> :func:`networkx_order`
  * ``from_networkx(G).edge_index`` column order          (:93)                             -> :func:`from_networkx_order`
"""
from __future__ import annotations

import warnings

import numpy as np


def _as_numpy_edge_index(edge_index) -> np.ndarray:
    if hasattr(edge_index, "detach"):
        edge_index = edge_index.detach().cpu().numpy()
    ei = np.asarray(edge_index, dtype=np.int64)
    if ei.ndim != 2 or ei.shape[0] != 2:
        raise ValueError("edge_index must have shape [2, E]")
    return ei


def undirected_csr(edge_index, num_nodes: int | None = None):
    """Sorted CSR of the symmetrised, self-loop-free, de-duplicated graph: ``(rowptr int32[n+1], col int32)``.

    ``n`` defaults to ``max index + 1`` — the reference's ``N = A.shape[0]`` (sdrf_cuda_bfc.py:29-30).
    """
    ei = _as_numpy_edge_index(edge_index)
    n = int(ei.max()) + 1 if ei.size else 0
    if num_nodes is not None:
        n = max(n, int(num_nodes))
    src = np.concatenate([ei[0], ei[1]])
    dst = np.concatenate([ei[1], ei[0]])
    keep = src != dst
    key = np.unique(src[keep] * max(n, 1) + dst[keep])
    src = key // max(n, 1)
    dst = key % max(n, 1)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(src, minlength=n), out=rowptr[1:])
    if rowptr[-1] >= 2**31:
        raise ValueError("graph too large for int32 CSR offsets")
    return rowptr.astype(np.int32), dst.astype(np.int32)


def networkx_order(edge_index, num_nodes: int):
    """Adjacency of ``to_networkx(data).to_undirected()`` in networkx insertion order.

    Returns ``(rowptr int32[n+1], order int32[nnz])``: ``order[rowptr[v]:rowptr[v+1]]`` lists the neighbours of
    ``v`` in the order ``G.neighbors(v)`` yields them.  A pair ``{u,v}`` enters both adjacency dicts when
    ``DiGraph.to_undirected`` first meets one of its directed edges, iterating sources in node order and, per
    source, successors in order of first appearance in ``edge_index``; so neighbours are ordered by that
    first-touch time.  Self-loops are dropped (with a warning): the reference keeps them in ``G`` but not in ``A``
    (sdrf_cuda_bfc.py:29 vs :31), a mismatch no shipped dataset exercises.
    """
    ei = _as_numpy_edge_index(edge_index)
    n = int(num_nodes)
    if ei.size and int(ei.max()) >= n:
        n = int(ei.max()) + 1
    u, v = ei[0], ei[1]
    loops = u == v
    if loops.any():
        warnings.warn("self-loops dropped from the rewiring graph (the reference keeps them in G but not in A)")
        u, v = u[~loops], v[~loops]
    m = u.size
    if m == 0:
        return np.zeros(n + 1, dtype=np.int32), np.zeros(0, dtype=np.int32)
    # first occurrence of every directed (u,v), in column order
    dkey = u * n + v
    _, first = np.unique(dkey, return_index=True)
    first.sort()
    u, v = u[first], v[first]
    # rank of v among the successors of u (order of first appearance)
    by_src = np.argsort(u, kind="stable")
    us = u[by_src]
    start = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(us, minlength=n), out=start[1:])
    rank = np.empty(u.size, dtype=np.int64)
    rank[by_src] = np.arange(u.size) - start[us]
    time = u * (int(rank.max()) + 1) + rank          # lexicographic (source, rank)
    # first touch of every undirected pair
    lo, hi = np.minimum(u, v), np.maximum(u, v)
    pkey = lo * n + hi
    order = np.lexsort((time, pkey))
    pk, tm = pkey[order], time[order]
    head = np.ones(pk.size, dtype=bool)
    head[1:] = pk[1:] != pk[:-1]
    pk, tm = pk[head], tm[head]
    a, b = pk // n, pk % n
    rows = np.concatenate([a, b])
    cols = np.concatenate([b, a])
    tms = np.concatenate([tm, tm])
    o = np.lexsort((tms, rows))
    rows, cols = rows[o], cols[o]
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=rowptr[1:])
    return rowptr.astype(np.int32), cols.astype(np.int32)


def from_networkx_order(rowptr: np.ndarray, order: np.ndarray) -> np.ndarray:
    """``from_networkx(G).edge_index`` (int64 ``[2, nnz]``) for adjacency lists in insertion order.

    PyG's ``from_networkx`` rebuilds the graph with ``nx.convert_node_labels_to_integers`` (iterating ``G.edges``,
    i.e. each undirected edge once from its first endpoint in node order) before listing the directed edges, so
    row ``w`` of the result holds the neighbours ``< w`` ascending, then the neighbours ``> w`` in insertion order.
    """
    rowptr = np.asarray(rowptr, dtype=np.int64)
    order = np.asarray(order, dtype=np.int64)
    n = rowptr.size - 1
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr))
    pos = np.arange(order.size, dtype=np.int64) - rowptr[rows]
    earlier = order < rows
    # sort key inside a row: earlier neighbours by id, later ones by insertion position (after all earlier ones)
    key = np.where(earlier, order, n + pos)
    o = np.lexsort((key, rows))
    return np.stack([rows[o], order[o]])

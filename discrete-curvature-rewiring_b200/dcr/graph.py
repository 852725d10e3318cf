"""Host-side graph set-up for the SDRF / BFC entry points (numpy; runs once per call, not in the hot loop).

Restates, vectorised, the third-party canonicalisation the reference relies on:
  * ``to_undirected`` + ``remove_self_loops``           (rewiring/sdrf_cuda_bfc.py:26-29)  -> :func:`undirected_csr`
  * ``to_networkx(data).to_undirected()`` insertion order (:31-33)                          -> :func:`networkx_order`
  * ``from_networkx(G).edge_index`` column order          (:93)                             -> :func:`from_networkx_order`
  * ``to_networkx(data)`` kept as a ``DiGraph`` (is_undirected=False, :31, :48-49)          -> :func:`digraph_order`,
    :func:`from_digraph_order`
  * ``to_networkx(data, node_attrs=['x'], to_undirected=True)`` (rewiring/sdrf_no_cuda.py:19) -> :func:`classical_order`
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


def networkx_order(edge_index, num_nodes: int, keep_self_loops: bool = False):
    """Adjacency of ``to_networkx(data).to_undirected()`` in networkx insertion order.

    Returns ``(rowptr int32[n+1], order int32[nnz])``: ``order[rowptr[v]:rowptr[v+1]]`` lists the neighbours of
    ``v`` in the order ``G.neighbors(v)`` yields them.  A pair ``{u,v}`` enters both adjacency dicts when
    ``DiGraph.to_undirected`` first meets one of its directed edges, iterating sources in node order and, per
    source, successors in order of first appearance in ``edge_index``; so neighbours are ordered by that
    first-touch time.  Self-loops: the reference keeps them in ``G`` but not in ``A`` (sdrf_cuda_bfc.py:29 vs :31), so a
    node with a self-loop lists ITSELF among its neighbours (and then once more at the end of its candidate list, :45-46).
    ``keep_self_loops=True`` reproduces that — ``v`` appears in its own list at its insertion position (the BFC loop's
    arena keeps such an entry in the insertion-order row only); otherwise they are dropped with a warning.
    """
    ei = _as_numpy_edge_index(edge_index)
    n = int(num_nodes)
    if ei.size and int(ei.max()) >= n:
        n = int(ei.max()) + 1
    u, v = ei[0], ei[1]
    loops = u == v
    if loops.any() and not keep_self_loops:
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
    two = a != b                                     # a self-loop enters ONE adjacency dict, once
    rows = np.concatenate([a, b[two]])
    cols = np.concatenate([b, a[two]])
    tms = np.concatenate([tm, tm[two]])
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


def _first_occurrences(u: np.ndarray, v: np.ndarray, n: int):
    """Columns of ``(u, v)`` without repeats of a directed pair, in order of first appearance."""
    _, first = np.unique(u * n + v, return_index=True)
    first.sort()
    return u[first], v[first]


def _rows_in_time_order(rows: np.ndarray, cols: np.ndarray, n: int):
    """CSR ``(rowptr, cols)`` with each row's entries in the order they appear in the input."""
    o = np.argsort(rows, kind="stable")
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=rowptr[1:])
    return rowptr.astype(np.int32), cols[o].astype(np.int32)


def digraph_order(edge_index, num_nodes: int, keep_self_loops: bool = False):
    """Successor and predecessor lists of ``to_networkx(data)`` — a ``DiGraph`` filled by ``add_edge(u, v)`` in column
    order (rewiring/sdrf_cuda_bfc.py:31; is_undirected=False keeps it directed) — in networkx insertion order:
    ``(succ_rowptr, succ_order, pred_rowptr, pred_order)``.  ``G.successors(x)`` / ``G.predecessors(y)`` (:48-49) yield
    exactly these orders.  Self-loops: kept with ``keep_self_loops=True`` (``v`` then sits in its own successor AND
    predecessor list, like in the reference's ``G``; ``A`` has no diagonal), otherwise dropped with a warning;
    a repeated directed pair raises: ``to_dense_adj`` would sum it into a weight 2 (:29), which the 0/1 kernels do not
    model."""
    ei = _as_numpy_edge_index(edge_index)
    n = int(num_nodes)
    if ei.size and int(ei.max()) >= n:
        n = int(ei.max()) + 1
    u, v = ei[0], ei[1]
    loops = u == v
    if loops.any() and not keep_self_loops:
        warnings.warn("self-loops dropped from the rewiring graph (the reference keeps them in G but not in A)")
        u, v = u[~loops], v[~loops]
    if u.size and np.unique(u * n + v).size != u.size:
        raise NotImplementedError("directed SDRF: repeated directed edges would make A weighted (to_dense_adj sums "
                                  "them); only 0/1 adjacency is supported")
    s_rp, s_ord = _rows_in_time_order(u, v, n)
    p_rp, p_ord = _rows_in_time_order(v, u, n)
    return s_rp, s_ord, p_rp, p_ord


def from_digraph_order(rowptr: np.ndarray, order: np.ndarray) -> np.ndarray:
    """``from_networkx(G).edge_index`` for a ``DiGraph``: ``convert_node_labels_to_integers`` re-adds the edges in
    ``G.edges`` order — sources in node order, successors in insertion order — which is the order of the result."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    rows = np.repeat(np.arange(rowptr.size - 1, dtype=np.int64), np.diff(rowptr))
    return np.stack([rows, np.asarray(order, dtype=np.int64)])


def classical_order(edge_index, num_nodes: int, keep_self_loops: bool = False):
    """Adjacency of ``to_networkx(data, node_attrs=['x'], to_undirected=True)`` (rewiring/sdrf_no_cuda.py:19) in networkx
    insertion order: PyG 2.0.3 skips every column with ``v > u`` and calls ``add_edge(u, v)`` for the others, in column
    order — an edge listed only as ``(u, v)`` with ``u < v`` is therefore lost, like in the reference.  Self-loops:
    kept with ``keep_self_loops=True`` (``u`` then sits in its own list where ``add_edge(u, u)`` happened), otherwise
    dropped with a warning.  Returns ``(rowptr int32[n+1], order int32[nnz])``."""
    ei = _as_numpy_edge_index(edge_index)
    n = int(num_nodes)
    if ei.size and int(ei.max()) >= n:
        n = int(ei.max()) + 1
    u, v = ei[0], ei[1]
    if (u == v).any() and not keep_self_loops:
        warnings.warn("self-loops dropped from the rewiring graph (the reference's sdrf_no_cuda would keep them)")
    keep = (v <= u) if keep_self_loops else (v < u)       # a column (u, u) passes PyG's `v > u` filter: add_edge(u, u)
    u, v = u[keep], v[keep]
    if u.size == 0:
        return np.zeros(n + 1, dtype=np.int32), np.zeros(0, dtype=np.int32)
    u, v = _first_occurrences(u, v, n)
    t = np.arange(u.size, dtype=np.int64)
    two = u != v                                           # a loop enters one adjacency dict, once
    rows = np.concatenate([u, v[two]])
    cols = np.concatenate([v, u[two]])
    tm = np.concatenate([t, t[two]])
    o = np.lexsort((tm, rows))
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=rowptr[1:])
    return rowptr.astype(np.int32), cols[o].astype(np.int32)

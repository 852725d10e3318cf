"""Oracle (TEST INFRASTRUCTURE ONLY): CPU restatement of the paper-flavour Balanced Forman curvature.

Follows ``/root/reference/curvature/bfc_naive.py``: ``bfc_edge`` :7-40 and ``bfc`` :43-52, on plain Python
sets (no networkx / scipy): the reference's sparse row products ``A[k] @ (A[v2] - A[v1]∘A[v2])ᵀ`` (:36-37) are
``|N(k) ∩ (N(v2) \\ N(v1))|`` for a 0/1 adjacency matrix.  Integer outputs (degrees, #triangles, #squares_1,
#squares_2, gamma) are the bit-exact parity targets; the value is evaluated left to right in Python floats
exactly as :31-32 / :39-40 do.

Parity pinning: ``tests/golden/paper_*.npz`` hold outputs of the unmodified ``bfc_naive.bfc_edge`` (run in the
build container with an ``nx.adj_matrix`` shim, see ``tests/golden/generate_golden.py``) and SURVEY.md App. G.
"""
from __future__ import annotations

import numpy as np


def adjacency_sets(edge_index: np.ndarray, n: int) -> list[set]:
    adj = [set() for _ in range(n)]
    for u, v in zip(np.asarray(edge_index[0]).tolist(), np.asarray(edge_index[1]).tolist()):
        if u != v:
            adj[u].add(v)
            adj[v].add(u)
    return adj


def bfc_edge_fields(adj: list[set], v1: int, v2: int):
    """``(deg1, deg2, triangles, squares_1, squares_2, gamma, value)`` for edge ``(v1, v2)``.

    ``gamma`` is 0 where the reference never computes it (either squares set empty, or ``deg_min == 1``).
    ``value`` is the Python ``int`` 0 when ``deg_min == 1`` (``bfc_naive.py:18-19``), else a Python float.
    """
    S1_1 = adj[v1]
    S1_2 = adj[v2]
    deg1 = len(S1_1)                              # :15-16  (G.degree; no self-loops in scope)
    deg2 = len(S1_2)
    deg_min = min(deg1, deg2)                     # :17
    triangles = S1_1 & S1_2                       # :25
    if deg_min == 1:                              # :18-19
        return deg1, deg2, len(triangles), 0, 0, 0, 0
    deg_max = max(deg1, deg2)                     # :20
    squares_1 = {k for k in S1_1 - S1_2 if k != v2 and (adj[k] & S1_2) - (S1_1 | {v1})}   # :26-27
    squares_2 = {k for k in S1_2 - S1_1 if k != v1 and (adj[k] & S1_1) - (S1_2 | {v2})}   # :28-29
    tri = len(triangles)
    if len(squares_1) == 0 or len(squares_2) == 0:                                        # :30-32
        value = 2 / deg1 + 2 / deg2 - 2 + 2 * tri / deg_max + tri / deg_min
        return deg1, deg2, tri, len(squares_1), len(squares_2), 0, value
    only2 = S1_2 - S1_1                           # support of A[v2] - A[v1]∘A[v2]      :36
    only1 = S1_1 - S1_2                           # support of A[v1] - A[v2]∘A[v1]      :37
    gamma = max(max(len(adj[k] & only2) - 1 for k in squares_1),
                max(len(adj[k] & only1) - 1 for k in squares_2))
    value = 2 / deg1 + 2 / deg2 - 2 + 2 * tri / deg_max + tri / deg_min + 1 / gamma / deg_max * (
        len(squares_1) + len(squares_2))          # :39-40
    return deg1, deg2, tri, len(squares_1), len(squares_2), int(gamma), value


def bfc_paper(edge_index: np.ndarray, n: int, edges: np.ndarray | None = None) -> dict:
    """Per undirected edge ``(i<j)`` (or the given ``edges`` ``[E,2]``) arrays of the fields above (:43-52)."""
    adj = adjacency_sets(edge_index, n)
    if edges is None:
        src = np.asarray(edge_index[0])
        dst = np.asarray(edge_index[1])
        m = src < dst
        edges = np.stack([src[m], dst[m]], axis=1)
        edges = edges[np.lexsort((edges[:, 1], edges[:, 0]))]
    out = np.zeros((len(edges), 6), dtype=np.int64)
    val = np.zeros(len(edges), dtype=np.float64)
    for e, (a, b) in enumerate(np.asarray(edges).tolist()):
        f = bfc_edge_fields(adj, a, b)
        out[e] = f[:6]
        val[e] = float(f[6])
    return {"edges": np.asarray(edges, dtype=np.int64), "deg_i": out[:, 0], "deg_j": out[:, 1], "tri": out[:, 2],
            "sq_i": out[:, 3], "sq_j": out[:, 4], "gamma": out[:, 5], "bfc": val}

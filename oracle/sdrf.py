"""Oracle (TEST INFRASTRUCTURE ONLY): CPU restatement of the BFC-SDRF rewiring loop.

Follows ``/root/reference/rewiring/sdrf_cuda_bfc.py:14-93`` statement by statement on dense numpy matrices,
calling the dense kernel restatements of ``oracle/cuda_flavour.py`` where the reference calls
``balanced_forman_curvature`` (:39) and ``balanced_forman_post_delta`` (:57), ``utils/softmax.py:4-10`` for the
probabilities (:66) and the documented algorithm of ``numpy.random.choice`` for the draw
(:64-68).  Third-party behaviour restated here (sources are not under /root/reference; versions pinned in
the reference's requirements.txt: torch-geometric 2.0.3, networkx 2.6.3, numpy 1.21.5):
  * ``to_undirected`` / ``remove_self_loops`` / ``to_dense_adj``            (:28-29)   -> :func:`dense_adjacency`
  * ``to_networkx(data).to_undirected()`` adjacency *insertion order*     (:31-33)   -> :func:`networkx_adjacency`
  * ``from_networkx(G).edge_index``                                       (:93)      -> :func:`edge_index_from_adjacency`

The only liberty taken is speed: ``A2 = A @ A`` may be maintained by exact rank-one integer updates
(``incremental_a2=True``) instead of being recomputed twice per iteration; ``verify_a2_every`` recomputes the
product and asserts equality.

Parity pinning: ``tests/golden/sdrf_*.npz`` hold the add/remove sequences of the UNMODIFIED
``rewiring.rewire.rewire(..., 'bfc', ...)`` run under ``NUMBA_ENABLE_CUDASIM=1`` (rounding model ``"sim32"``).
"""
from __future__ import annotations

import numpy as np

from .cuda_flavour import bfc_cuda_dense, closing_value, post_delta_dense, F32, F64


# ----------------------------------------------------------------------------------------------------------
# utils/softmax.py:4-10 and the np.random.choice draw (sdrf_cuda_bfc.py:64-68)
# ----------------------------------------------------------------------------------------------------------
def softmax(a: np.ndarray, tau=1) -> np.ndarray:
    """``utils/softmax.py:4-10`` — one-hot at the first argmax for ``tau == inf``, else ``exp(a*tau)/sum``."""
    if tau == float("inf"):
        r = np.zeros(len(a))
        r[np.argmax(a)] = 1
        return r
    exp_a = np.exp(a * tau)
    return exp_a / exp_a.sum()


def choice_index(p: np.ndarray, u: float) -> int:
    """``np.random.choice(range(n), p=p)`` for the uniform ``u`` it would draw.

    Raises ``ValueError`` like numpy does for NaN / negative probabilities or a sum off by more than
    ``sqrt(eps)``.
    """
    p = np.asarray(p, dtype=np.float64)
    if np.isnan(p).any():
        raise ValueError("probabilities contain NaN")
    if (p < 0).any():
        raise ValueError("probabilities are not non-negative")
    if abs(p.sum() - 1.0) > np.sqrt(np.finfo(np.float64).eps):
        raise ValueError("probabilities do not sum to 1")
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return int(cdf.searchsorted(u, side="right"))


# ----------------------------------------------------------------------------------------------------------
# graph set-up / tear-down (third-party semantics)
# ----------------------------------------------------------------------------------------------------------
def dense_adjacency(edge_index: np.ndarray, is_undirected: bool = True) -> np.ndarray:
    """``to_dense_adj(remove_self_loops(to_undirected(edge_index))[0])[0]`` (sdrf_cuda_bfc.py:26-29)."""
    ei = np.asarray(edge_index, dtype=np.int64)
    n = int(ei.max()) + 1 if ei.size else 0
    if is_undirected:
        ei = np.concatenate([ei, ei[::-1]], axis=1)
        key = np.unique(ei[0] * n + ei[1])       # coalesce: sort + drop duplicates
        ei = np.stack([key // n, key % n])
    ei = ei[:, ei[0] != ei[1]]
    A = np.zeros((n, n), dtype=F32)
    np.add.at(A, (ei[0], ei[1]), F32(1))        # duplicates (directed input only) sum
    return A


def networkx_adjacency(edge_index: np.ndarray, num_nodes: int) -> list[dict]:
    """Adjacency dicts, in networkx insertion order, of ``to_networkx(data).to_undirected()`` (:31-33)."""
    succ = [dict() for _ in range(num_nodes)]
    for u, v in zip(np.asarray(edge_index[0]).tolist(), np.asarray(edge_index[1]).tolist()):
        succ[u][v] = None                        # DiGraph.add_edge in column order
    adj = [dict() for _ in range(num_nodes)]
    for u in range(num_nodes):                   # Graph.add_edges_from over DiGraph._adj
        for v in succ[u]:
            adj[u][v] = None
            adj[v][u] = None
    return adj


def edge_index_from_adjacency(adj: list[dict]) -> np.ndarray:
    """``from_networkx(G).edge_index`` (:93).

    PyG 2.0.3 ``from_networkx`` first calls ``nx.convert_node_labels_to_integers(G)``, which REBUILDS the graph
    by iterating ``G.edges`` (each undirected edge once, reported from its first endpoint in node order) and
    only then takes ``list(G.to_directed().edges)``.  So the adjacency order of the copy — and hence the column
    order of the result — is: neighbours that precede ``w`` in node order (ascending), then the remaining
    neighbours in ``G``'s insertion order.
    """
    n = len(adj)
    copy = [dict() for _ in range(n)]
    seen = set()
    for u in range(n):                           # Graph.edges: skip neighbours already visited as `u`
        for v in adj[u]:
            if v not in seen:
                copy[u][v] = None
                copy[v][u] = None
        seen.add(u)
    src, dst = [], []
    for u, nbrs in enumerate(copy):              # DiGraph built by to_directed(): same per-node order
        for v in nbrs:
            src.append(u)
            dst.append(v)
    return np.array([src, dst], dtype=np.int64).reshape(2, -1)


def networkx_digraph(edge_index: np.ndarray, num_nodes: int):
    """``(succ, pred)`` adjacency dicts, in networkx insertion order, of ``to_networkx(data)`` — a ``DiGraph`` built
    by ``add_edge(u, v)`` in column order of ``edge_index`` (:31, directed mode keeps it as is)."""
    succ = [dict() for _ in range(num_nodes)]
    pred = [dict() for _ in range(num_nodes)]
    for u, v in zip(np.asarray(edge_index[0]).tolist(), np.asarray(edge_index[1]).tolist()):
        succ[u][v] = None
        pred[v][u] = None
    return succ, pred


def edge_index_from_digraph(succ: list[dict]) -> np.ndarray:
    """``from_networkx(G).edge_index`` for a ``DiGraph`` (:93): ``convert_node_labels_to_integers`` re-adds the edges
    in ``G.edges`` order (node order, successors in insertion order), which is also the order of the result."""
    src, dst = [], []
    for u, nbrs in enumerate(succ):
        for v in nbrs:
            src.append(u)
            dst.append(v)
    return np.array([src, dst], dtype=np.int64).reshape(2, -1)


def first_argmin(C: np.ndarray) -> int:
    """``C.argmin().item()`` (:40): first row-major occurrence; ``-0.0 == 0.0``."""
    return int(np.argmin(C))


def first_argmax(C: np.ndarray) -> int:
    return int(np.argmax(C))


# ----------------------------------------------------------------------------------------------------------
# the loop
# ----------------------------------------------------------------------------------------------------------
def sdrf_oracle(edge_index: np.ndarray, num_nodes: int, loops: int, remove_edges: bool, removal_bound: float,
                tau, uniforms: np.ndarray | None = None, rounding: str = "compiled",
                incremental_a2: bool = True, verify_a2_every: int = 0, is_undirected: bool = True):
    """Returns ``(edge_index_out, log)``.

    ``log`` is a list with one dict per executed iteration:
    ``{"x","y","n_candidates","k","l","choice","removed": (a,b) or None, "improvements": fp64 array}``
    (``k = l = choice = -1`` when nothing was added).  ``uniforms[t]`` is the t-th double ``np.random`` would
    have produced (one is consumed per iteration that has candidates, also for ``tau == inf``).
    ``is_undirected=False`` (no caller of the reference uses it, rewire.py:10) restates the directed branches
    (:47-49, :72-73, :87-88): candidates from the successors of ``x`` and the predecessors of ``y``, one directed
    entry added / removed, ``G`` a ``DiGraph``.
    """
    A = dense_adjacency(edge_index, is_undirected)              # :26-29
    N = A.shape[0]                                              # :30
    if is_undirected:
        adj = networkx_adjacency(edge_index, max(num_nodes, N)) # :31-33
        pred = None
    else:
        adj, pred = networkx_digraph(edge_index, max(num_nodes, N))   # :31 (adj = successors)
    A2 = (A @ A).astype(F32)
    n_draws = 0
    log = []

    def toggle(a, b, val):
        nonlocal A2
        A[a, b] = val
        if is_undirected:
            A[b, a] = val
        if not incremental_a2:
            A2 = (A @ A).astype(F32)

    for it in range(loops):                                     # :37
        can_add = True                                          # :38
        if incremental_a2 and verify_a2_every and it % verify_a2_every == 0:
            assert np.array_equal(A2, (A @ A).astype(F32))
        res = _bfc_with_a2(A, A2, rounding)                     # :39
        C = res
        ix_min = first_argmin(C)                                # :40
        x, y = ix_min // N, ix_min % N                          # :41-42
        x_neighbors = list(adj[x]) + [x]                        # :45 / :48 (successors)
        y_neighbors = list(adj[y] if is_undirected else pred[y]) + [y]   # :46 / :49 (predecessors)
        candidates = [(i, j) for i in x_neighbors for j in y_neighbors
                      if (i != j) and (j not in adj[i])]        # :50-54
        rec = {"x": x, "y": y, "n_candidates": len(candidates), "k": -1, "l": -1, "choice": -1,
               "removed": None, "improvements": np.zeros(0)}
        stop = False
        if len(candidates):                                     # :56
            D = post_delta_dense(A, x, y, x_neighbors, y_neighbors, rounding, A2=A2)   # :57
            cxy = C[x, y]
            improvements = []
            for (i, j) in candidates:                           # :59-62  fp32 subtraction, widened by .item()
                improvements.append(float(F32(D[x_neighbors.index(i), y_neighbors.index(j)] - cxy)))
            improvements = np.array(improvements)
            p = softmax(improvements, tau=tau)                  # :66
            if uniforms is None:
                raise ValueError("uniforms are required (one per iteration with candidates)")
            choice = choice_index(p, float(uniforms[n_draws]))  # :64-68
            n_draws += 1
            k, l = candidates[choice]
            adj[k][l] = None                                    # :69  G.add_edge(k, l)
            if is_undirected:
                adj[l][k] = None
            else:
                pred[l][k] = None
            if incremental_a2:
                (_a2_toggle if is_undirected else _a2_toggle_directed)(A, A2, k, l, +1)
            toggle(k, l, F32(1))                                # :70-71
            rec.update(k=k, l=l, choice=choice, improvements=improvements)
        else:
            can_add = False                                     # :75
            if not remove_edges:                                # :76-77
                stop = True
        if remove_edges and not stop:                           # :79
            ix_max = first_argmax(C)                            # :80   (C is the pre-add matrix)
            xr, yr = ix_max // N, ix_max % N                    # :81-82
            if C[xr, yr] > removal_bound:                       # :83
                if yr not in adj[xr]:
                    raise KeyError(f"The edge {xr}-{yr} is not in the graph")   # networkx.NetworkXError
                del adj[xr][yr]                                 # :84  G.remove_edge
                if is_undirected:
                    if xr != yr:
                        del adj[yr][xr]
                else:
                    del pred[yr][xr]
                if incremental_a2:
                    (_a2_toggle if is_undirected else _a2_toggle_directed)(A, A2, xr, yr, -1)
                toggle(xr, yr, F32(0))                          # :85-86
                rec["removed"] = (xr, yr)
            else:
                if can_add is False:                            # :89-91
                    stop = True
        log.append(rec)
        if stop:
            break
    return (edge_index_from_adjacency(adj) if is_undirected else edge_index_from_digraph(adj)), log   # :93


def _a2_toggle(A: np.ndarray, A2: np.ndarray, k: int, l: int, sign: int) -> None:
    """Exact update of ``A2 = A @ A`` for ``A[k,l] = A[l,k] += sign`` (call BEFORE changing ``A``).

    ``(A+Δ)² = A² + AΔ + ΔA + Δ²`` with ``Δ = sign·(e_k e_lᵀ + e_l e_kᵀ)``.
    """
    s = F32(sign)
    A2[:, l] += s * A[:, k]
    A2[:, k] += s * A[:, l]
    A2[k, :] += s * A[l, :]
    A2[l, :] += s * A[k, :]
    A2[k, k] += F32(1)
    A2[l, l] += F32(1)


def _a2_toggle_directed(A: np.ndarray, A2: np.ndarray, k: int, l: int, sign: int) -> None:
    """Exact update of ``A2 = A @ A`` for the single directed entry ``A[k,l] += sign`` (call BEFORE changing ``A``):
    ``(A+Δ)² = A² + AΔ + ΔA + Δ²`` with ``Δ = sign·e_k e_lᵀ``; ``Δ² = 0`` for ``k != l``."""
    s = F32(sign)
    A2[:, l] += s * A[:, k]
    A2[k, :] += s * A[l, :]
    if k == l:
        A2[k, k] += F32(1)


def _bfc_with_a2(A: np.ndarray, A2: np.ndarray, rounding: str) -> np.ndarray:
    """:func:`oracle.cuda_flavour.bfc_cuda_dense` with a caller-supplied ``A2`` (same statements)."""
    N = A.shape[0]
    d_in = A.sum(axis=0, dtype=F32)
    d_out = A.sum(axis=1, dtype=F32)
    C = np.zeros((N, N), dtype=F32)
    AT = A.T
    A2T = A2.T
    for i, j in zip(*np.nonzero(A)):
        if d_in[i] > d_out[j]:
            d_max, d_min = d_in[i], d_out[j]
        else:
            d_max, d_min = d_out[j], d_in[i]
        if d_max * d_min == 0:
            continue
        aij = A[i, j]
        t1 = AT[j] * (A2[i] - A[i]) * aij
        t2 = A[i] * (A2T[j] - AT[j]) * aij
        p1 = t1 > 0
        p2 = t2 > 0
        s_ij = int(p1.sum()) + int(p2.sum())
        l_ij = 0.0
        if s_ij:
            l_ij = float(max(t1.max(), t2.max()))
        C[i, j] = closing_value(d_max, d_min, A2[i, j], aij, s_ij, l_ij, rounding)
    return C

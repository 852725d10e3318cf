"""Oracle (TEST INFRASTRUCTURE ONLY): CPU restatement of the classical-curvature SDRF loop.

Follows ``/root/reference/rewiring/sdrf_no_cuda.py:9-68`` statement by statement on plain adjacency dicts, with
``/root/reference/curvature/classical_curvatures.py:6-46`` for the three curvatures and ``utils/softmax.py`` +
the ``np.random.choice`` algorithm (``oracle/sdrf.py``) for the draw.  Third-party behaviour restated here
(torch-geometric 2.0.3, networkx 2.6.3; sources not under /root/reference):
  * ``to_networkx(data, node_attrs=['x'], to_undirected=True)`` (:19): an ``nx.Graph`` that receives ``add_edge(u, v)``
    in column order of ``edge_index`` for the columns with ``v <= u`` ONLY — a column ``(u, v)`` with ``v > u`` is
    skipped, so an input that lists an edge in one direction ``u < v`` only loses it (the reference's behaviour);
  * ``G.edges`` (:27, :59-61): every edge once, reported from its first endpoint in node order (nodes were added as
    ``range(num_nodes)``), neighbours in adjacency insertion order; ``min`` / ``max`` return the FIRST extreme;
  * ``from_networkx(G).edge_index`` (:68): ``oracle.sdrf.edge_index_from_adjacency``.

Parity pinning: ``tests/golden/sdrf_classical_seq.npz`` holds add / remove sequences and outputs of the UNMODIFIED
``rewiring.rewire.rewire(data, curv_type, ...)`` for the three curvature types (``tests/golden/generate_golden.py``).
"""
from __future__ import annotations

import numpy as np

from .sdrf import choice_index, edge_index_from_adjacency, softmax


def networkx_adjacency_classical(edge_index: np.ndarray, num_nodes: int) -> list[dict]:
    """Adjacency dicts of ``to_networkx(data, to_undirected=True)`` (sdrf_no_cuda.py:19), self-loops included."""
    adj = [dict() for _ in range(num_nodes)]
    for u, v in zip(np.asarray(edge_index[0]).tolist(), np.asarray(edge_index[1]).tolist()):
        if v > u:
            continue
        adj[u][v] = None
        adj[v][u] = None
    return adj


def _degree(adj, v):
    return len(adj[v]) + (1 if v in adj[v] else 0)      # networkx counts a self-loop twice


def curvature_edge(adj, e, curv_type: str) -> int:
    """``compute_curvature_edge`` (classical_curvatures.py:6-28)."""
    v1, v2 = e
    if curv_type == "1d":
        return 4 - _degree(adj, v1) - _degree(adj, v2)
    tri = len(set(adj[v1]) & set(adj[v2]))
    if curv_type == "augmented":
        return 4 - _degree(adj, v1) - _degree(adj, v2) + 3 * tri
    if curv_type == "haantjes":
        return tri
    raise Exception(f"Method {curv_type} not available.")


def graph_edges(adj):
    """``G.edges`` of an ``nx.Graph`` whose nodes are ``0..n-1`` in order."""
    seen = set()
    for u in range(len(adj)):
        for v in adj[u]:
            if v not in seen:
                yield (u, v)
        seen.add(u)


def sdrf_classical_oracle(edge_index: np.ndarray, num_nodes: int, curv_type: str, loops: int, remove_edges: bool,
                          removal_bound: float, tau, uniforms: np.ndarray | None = None):
    """Returns ``(edge_index_out, log)``; ``log`` has one dict per executed iteration:
    ``{"x","y","n_candidates","k","l","choice","removed","improvements"}``."""
    adj = networkx_adjacency_classical(edge_index, num_nodes)                     # :19
    n_draws = 0
    log = []
    for _ in range(loops):                                                        # :21
        can_add = True                                                            # :22
        edges = list(graph_edges(adj))
        curv = {e: curvature_edge(adj, e, curv_type) for e in edges}              # :23
        x, y = min(edges, key=lambda e: curv[e])                                  # :26 (ValueError on an empty graph)
        x_neighbors = list(adj[x]) + [x]                                          # :28-29
        y_neighbors = list(adj[y]) + [y]
        candidates = [sorted((i, j)) for i in x_neighbors for j in y_neighbors
                      if (i != j) and (j not in adj[i])]                          # :31-36
        rec = {"x": x, "y": y, "n_candidates": len(candidates), "k": -1, "l": -1, "choice": -1, "removed": None,
               "improvements": np.zeros(0)}
        stop = False
        k = l = None
        if len(candidates):                                                       # :38
            improvements = []
            for (i, j) in candidates:                                             # :40-45
                before = curvature_edge(adj, (x, y), curv_type)
                adj[i][j] = None
                adj[j][i] = None
                after = curvature_edge(adj, (x, y), curv_type)
                improvements.append(after - before)
                del adj[i][j]
                del adj[j][i]
            improvements = np.array(improvements)
            if uniforms is None:
                raise ValueError("uniforms are required (one per iteration with candidates)")
            choice = choice_index(softmax(improvements, tau=tau), float(uniforms[n_draws]))   # :48-50
            n_draws += 1
            k, l = sorted(candidates[choice])
            adj[k][l] = None                                                      # :51
            adj[l][k] = None
            rec.update(k=k, l=l, choice=choice, improvements=improvements.astype(np.float64))
        else:
            can_add = False                                                       # :53-55
            if not remove_edges:
                stop = True
        if remove_edges and not stop:                                             # :56
            now = list(graph_edges(adj))
            if len(candidates):                                                   # :58-61
                xr, yr = max([e for e in now if e != (k, l)], key=lambda e: curv[e])
            else:
                xr, yr = max(now, key=lambda e: curv[e])
            if curv[(xr, yr)] > removal_bound:                                    # :62-63
                del adj[xr][yr]
                if xr != yr:
                    del adj[yr][xr]
                rec["removed"] = (xr, yr)
            else:
                if can_add is False:                                              # :64-66
                    stop = True
        log.append(rec)
        if stop:
            break
    return edge_index_from_adjacency(adj), log                                    # :68

"""Seeded synthetic graphs of the shapes named in BASELINE.json.

There is no network on the build or GPU boxes, so every benchmark and parity graph is generated:
a power-law Chung-Lu graph (weights ``w_v ∝ (rank_v+1)^-alpha``), isolated nodes attached by one edge,
a triangle-closure pass that supplies ``p_tri`` of the edges (non-trivial #triangles / #4-cycles), and a
random relabelling, all driven by one ``numpy.random.default_rng(seed)``.  The result is an undirected simple
graph with node ids ``0..N-1`` returned as a sorted, symmetric ``edge_index`` (int64 ``[2, 2E]``) — the form
for which the reference's networkx adjacency order is ascending.
"""
from __future__ import annotations

import numpy as np

# name -> (N, E undirected, alpha, p_tri, seed); SURVEY.md §8d table.
SHAPES = {
    "cornell": (183, 295, 0.9, 0.1, 1183),
    "texas": (183, 309, 0.9, 0.1, 1184),
    "wisconsin": (251, 499, 0.9, 0.1, 1251),
    "cora": (2708, 5278, 0.5, 0.3, 2708),
    "squirrel": (5201, 198000, 0.7, 0.3, 5201),
    "arxiv": (169343, 1166243, 0.6, 0.1, 169343),
}

# name -> (loops, tau, removal_bound) from the reference's utils/hyperparams.py (values are benchmark inputs).
SDRF_PARAMS = {
    "cornell": (126, 145, 0.88),
    "texas": (89, 22, 1.64),
    "wisconsin": (136, 12, 7.95),
    "cora": (100, 163, 0.95),
    "squirrel": (1396, 436, 5.88),
}


def _unique_undirected(keys: np.ndarray) -> np.ndarray:
    return np.unique(keys)


def _csr_from_keys(keys: np.ndarray, n: int):
    u = keys // n
    v = keys % n
    src = np.concatenate([u, v])
    dst = np.concatenate([v, u])
    order = np.lexsort((dst, src))
    src = src[order]
    dst = dst[order]
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(src, minlength=n), out=rowptr[1:])
    return rowptr, dst


def chung_lu_graph(n: int, e: int, alpha: float, p_tri: float, seed: int) -> np.ndarray:
    """Return a sorted symmetric ``edge_index`` (int64 ``[2, 2e]``) with exactly ``e`` undirected edges."""
    if e > n * (n - 1) // 2:
        raise ValueError("more edges requested than a simple graph on n nodes can hold")
    rng = np.random.default_rng(seed)
    w = (np.arange(n, dtype=np.float64) + 1.0) ** (-alpha)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]

    def sample_nodes(k):
        return np.minimum(np.searchsorted(cdf, rng.random(k), side="right"), n - 1)

    def add_pairs(keys, a, b, limit):
        lo = np.minimum(a, b)
        hi = np.maximum(a, b)
        ok = lo != hi
        new = lo[ok] * n + hi[ok]
        # keep first occurrences in draw order so that the cut at `limit` is deterministic
        _, first = np.unique(new, return_index=True)
        new = new[np.sort(first)]
        new = new[~np.isin(new, keys)]
        room = limit - keys.size
        return np.concatenate([keys, new[:room]])

    keys = np.empty(0, dtype=np.int64)
    e_base = int(round(e * (1.0 - p_tri)))
    while keys.size < e_base:
        k = max(1024, int((e_base - keys.size) * 1.3))
        keys = add_pairs(keys, sample_nodes(k), sample_nodes(k), e_base)

    # attach isolated nodes with one edge each to a weight-sampled partner; if the edge budget cannot hold the
    # base edges plus one edge per isolated node, give back the most recently drawn base edges until it can
    def isolated(k):
        return np.flatnonzero(np.bincount(np.concatenate([k // n, k % n]), minlength=n) == 0)

    iso = isolated(keys)
    while keys.size + iso.size > e and keys.size > 0:
        keys = keys[: max(0, e - iso.size - max(1, iso.size // 8))]
        iso = isolated(keys)
    while iso.size and keys.size < e:
        keys = add_pairs(keys, iso, sample_nodes(iso.size), e)
        iso = isolated(keys)

    # triangle closure: pick a random directed edge (u,v) and a random neighbour w of v, add (u,w)
    stall = 0
    while keys.size < e:
        rowptr, col = _csr_from_keys(keys, n)
        k = max(1024, int((e - keys.size) * 1.5))
        pick = rng.integers(0, col.size, size=k)
        u = np.searchsorted(rowptr, pick, side="right") - 1
        v = col[pick]
        dv = rowptr[v + 1] - rowptr[v]
        wv = col[rowptr[v] + (rng.random(k) * dv).astype(np.int64)]
        before = keys.size
        keys = add_pairs(keys, u, wv, e)
        stall = stall + 1 if keys.size == before else 0
        if stall > 8:  # closure saturated (tiny dense graphs): fall back to Chung-Lu draws
            keys = add_pairs(keys, sample_nodes(k), sample_nodes(k), e)

    perm = rng.permutation(n)
    u = perm[keys // n]
    v = perm[keys % n]
    src = np.concatenate([u, v])
    dst = np.concatenate([v, u])
    order = np.lexsort((dst, src))
    return np.stack([src[order], dst[order]]).astype(np.int64)


def named_graph(name: str) -> tuple[np.ndarray, int]:
    """``(edge_index, num_nodes)`` for one of the BASELINE.json config shapes."""
    n, e, alpha, p_tri, seed = SHAPES[name]
    return chung_lu_graph(n, e, alpha, p_tri, seed), n


def csr_from_edge_index(edge_index: np.ndarray, n: int):
    """Sorted CSR (rowptr int64 ``[n+1]``, colidx int32) of a symmetric, sorted ``edge_index``."""
    src = np.asarray(edge_index[0])
    dst = np.asarray(edge_index[1])
    order = np.lexsort((dst, src))
    src = src[order]
    dst = dst[order]
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(src, minlength=n), out=rowptr[1:])
    return rowptr, dst.astype(np.int32)

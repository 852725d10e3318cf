"""Shared test helpers: graph builders and oracle/GPU comparison utilities."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sym_edge_index(pairs, n=None):
    e = np.array([(u, v) for u, v in pairs if u != v], dtype=np.int64).reshape(-1, 2)
    if e.size == 0:
        return np.zeros((2, 0), dtype=np.int64)
    nn = int(e.max()) + 1 if n is None else n
    src = np.concatenate([e[:, 0], e[:, 1]])
    dst = np.concatenate([e[:, 1], e[:, 0]])
    key = np.unique(src * nn + dst)
    return np.stack([key // nn, key % nn])


def gnp(n, p, seed):
    rng = np.random.default_rng(seed)
    iu = np.triu_indices(n, 1)
    m = rng.random(iu[0].size) < p
    return sym_edge_index(zip(iu[0][m].tolist(), iu[1][m].tolist()), n)


def dense_of(ei, n):
    A = np.zeros((n, n), dtype=np.float32)
    A[ei[0], ei[1]] = 1
    return A


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def toy_graphs():
    """Appendix G graphs as (edge_index, n), built without networkx."""
    def cyc(n):
        return [(i, (i + 1) % n) for i in range(n)]

    def clique(n, off=0):
        return [(off + i, off + j) for i in range(n) for j in range(i + 1, n)]

    g = {
        "path4": ([(0, 1), (1, 2), (2, 3)], 4),
        "star5": ([(0, i) for i in range(1, 5)], 5),
        "c3": (cyc(3), 3), "c4": (cyc(4), 4), "c5": (cyc(5), 5),
        "k4": (clique(4), 4), "k5": (clique(5), 5),
        "k33": ([(i, 3 + j) for i in range(3) for j in range(3)], 6),
        "grid3": ([(3 * r + c, 3 * r + c + 1) for r in range(3) for c in range(2)] +
                  [(3 * r + c, 3 * (r + 1) + c) for r in range(2) for c in range(3)], 9),
        "petersen": (cyc(5) + [(i, i + 5) for i in range(5)] + [(5 + i, 5 + (i + 2) % 5) for i in range(5)], 10),
        "cube": ([(a, a ^ (1 << b)) for a in range(8) for b in range(3) if a < a ^ (1 << b)], 8),
        "barbell41": (clique(4) + [(3, 4), (4, 5)] + clique(4, 5), 9),
    }
    return {k: (sym_edge_index(p, n), n) for k, (p, n) in g.items()}

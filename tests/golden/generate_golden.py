"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (``/root/reference`` is not present on the GPU box):

    python tests/golden/generate_golden.py

What is executed, all imported from ``/root/reference`` without edits:
  * ``curvature.bfc_naive.bfc_edge``                     (needs only an ``nx.adj_matrix`` shim under networkx>=3)
  * ``curvature.bfc_cuda.balanced_forman_curvature`` /
    ``balanced_forman_post_delta``                       under ``NUMBA_ENABLE_CUDASIM=1``; a proxy converts the
                                                         torch arguments of the kernel launch to numpy (with
                                                         torch arguments 0-d views alias under the simulator
                                                         and ``A2_x_y += …`` corrupts ``A2`` — SURVEY.md App. F.4)
  * ``rewiring.rewire.rewire(..., 'bfc', ...)``          with the ``torch_geometric`` stand-in of this repo and
                                                         ``torch.Tensor.cuda = identity``; the add/remove sequence
                                                         is captured by wrapping ``nx.Graph.add_edge/remove_edge``

The simulator evaluates the closing formula in fp32 (numpy>=2), i.e. the oracle's ``"sim32"`` rounding model;
the compiled kernel's fp64/two-rounding dataflow is pinned separately from its PTX.
  * ``rewiring.sdrf_cuda_bfc.sdrf_cuda_bfc(..., is_undirected=False)`` and ``rewire(data, '1d' | 'augmented' | 'haantjes', ...)``
    (``rewiring/sdrf_no_cuda.py``) the same way; all three loops also on inputs WITH self-loops

Outputs: ``paper_kat.npz``, ``paper_ints_kat.npz``, ``cuda_kat.npz``, ``sdrf_seq.npz``, ``sdrf_directed_seq.npz``,
``sdrf_classical_seq.npz``, ``sdrf_selfloop_seq.npz``, ``sdrf_directed_selfloop_seq.npz``, ``sdrf_classical_selfloop_seq.npz``
(a few hundred KB in total).
"""
from __future__ import annotations

import os
import sys
import time

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
os.environ.setdefault("NUMBA_DISABLE_PERFORMANCE_WARNINGS", "1")

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(REPO, "discrete-curvature-rewiring_b200")
REFERENCE = "/root/reference"
sys.path.insert(0, os.path.join(PKG, "standin"))
sys.path.insert(0, REFERENCE)
sys.path.insert(1, PKG)        # for dcr.synth only (the reference's own packages stay first)

import networkx as nx  # noqa: E402
import numpy as np  # noqa: E402
import scipy.sparse  # noqa: E402
import torch  # noqa: E402

if not hasattr(nx, "adj_matrix"):
    nx.adj_matrix = lambda G: scipy.sparse.csr_matrix(nx.adjacency_matrix(G))
torch.Tensor.cuda = lambda self, *a, **k: self

import curvature.bfc_cuda as ref_cuda  # noqa: E402
import curvature.bfc_naive as ref_naive  # noqa: E402
from torch_geometric.data import Data  # noqa: E402


class _NumpyArgProxy:
    """Forwards ``kernel[grid, block](*args)`` with torch tensors replaced by numpy views / scalars."""

    def __init__(self, kernel):
        self.kernel = kernel

    def __getitem__(self, cfg):
        launch = self.kernel[cfg]

        def call(*args):
            conv = []
            for a in args:
                if torch.is_tensor(a):
                    conv.append(np.float32(a.item()) if a.dim() == 0 else a.numpy())
                else:
                    conv.append(a)
            return launch(*conv)

        return call


ref_cuda._balanced_forman_curvature = _NumpyArgProxy(ref_cuda._balanced_forman_curvature)
ref_cuda._balanced_forman_post_delta = _NumpyArgProxy(ref_cuda._balanced_forman_post_delta)

import rewiring.sdrf_cuda_bfc as ref_sdrf  # noqa: E402
from rewiring.rewire import rewire as ref_rewire  # noqa: E402

ref_sdrf.tqdm = lambda it, *a, **k: it


def relabel(G):
    return nx.convert_node_labels_to_integers(G, ordering="sorted")


def toy_graphs():
    gs = {
        "path4": nx.path_graph(4),
        "star5": nx.star_graph(4),
        "c3": nx.cycle_graph(3),
        "c4": nx.cycle_graph(4),
        "c5": nx.cycle_graph(5),
        "k4": nx.complete_graph(4),
        "k5": nx.complete_graph(5),
        "k33": nx.complete_bipartite_graph(3, 3),
        "grid3": relabel(nx.grid_2d_graph(3, 3)),
        "petersen": nx.petersen_graph(),
        "cube": relabel(nx.hypercube_graph(3)),
        "barbell41": nx.barbell_graph(4, 1),
        "wheel7": nx.wheel_graph(7),
        "lollipop": nx.lollipop_graph(5, 3),
    }
    return gs


def sorted_symmetric_edge_index(G):
    e = np.array([(u, v) for u, v in G.edges() if u != v], dtype=np.int64).reshape(-1, 2)
    src = np.concatenate([e[:, 0], e[:, 1]])
    dst = np.concatenate([e[:, 1], e[:, 0]])
    order = np.lexsort((dst, src))
    return np.stack([src[order], dst[order]])


def graph_from_edge_index(ei, n):
    G = nx.Graph()
    G.add_nodes_from(range(n))
    G.add_edges_from(zip(ei[0].tolist(), ei[1].tolist()))
    return G


def gen_paper(out):
    from dcr.synth import named_graph

    graphs = {k: (sorted_symmetric_edge_index(g), g.number_of_nodes()) for k, g in toy_graphs().items()}
    for seed in range(24):
        n = 12 + seed
        g = nx.gnp_random_graph(n, 0.12 + 0.015 * (seed % 12), seed=seed)
        graphs[f"gnp{seed}"] = (sorted_symmetric_edge_index(g), n)
    for name in ("cornell", "wisconsin"):
        graphs[name] = named_graph(name)
    pack = {"names": np.array(sorted(graphs))}
    for name, (ei, n) in graphs.items():
        G = graph_from_edge_index(ei, n)
        m = ei[0] < ei[1]
        edges = np.stack([ei[0][m], ei[1][m]], axis=1)
        vals = np.array([float(ref_naive.bfc_edge(G, int(a), int(b))) for a, b in edges], dtype=np.float64)
        pack[f"{name}/edge_index"] = ei
        pack[f"{name}/n"] = np.int64(n)
        pack[f"{name}/edges"] = edges
        pack[f"{name}/bfc"] = vals
    np.savez_compressed(out, **pack)
    print("paper_kat:", len(graphs), "graphs")


def _bfc_edge_with_locals(G, v1, v2):
    """Run the UNMODIFIED ``bfc_edge`` and return ``(value, locals at return)`` — the integer quantities the paper
    flavour must match bit-exactly (``len(triangles)``, ``len(squares_1)``, ``len(squares_2)``, ``gamma``) are locals
    of that function, read from its frame by a profile hook."""
    grabbed = {}
    code = ref_naive.bfc_edge.__code__

    def hook(frame, event, arg):
        if event == "return" and frame.f_code is code:
            grabbed.update(frame.f_locals)

    sys.setprofile(hook)
    try:
        val = ref_naive.bfc_edge(G, v1, v2)
    finally:
        sys.setprofile(None)
    return val, grabbed


def gen_paper_ints(out):
    """Integer fields of the paper flavour from the unmodified reference: per edge (deg1, deg2, #triangles, #squares_1,
    #squares_2, gamma) with gamma = 0 where the reference never computes it (early returns at bfc_naive.py:18-19,30-32)."""
    from dcr.synth import named_graph

    graphs = {k: (sorted_symmetric_edge_index(g), g.number_of_nodes()) for k, g in toy_graphs().items()}
    for seed in range(24):
        n = 12 + seed
        g = nx.gnp_random_graph(n, 0.12 + 0.015 * (seed % 12), seed=seed)
        graphs[f"gnp{seed}"] = (sorted_symmetric_edge_index(g), n)
    for seed in range(6):                       # denser graphs: every edge reaches the gamma branch
        n = 30 + 4 * seed
        g = nx.gnp_random_graph(n, 0.3 + 0.05 * seed, seed=300 + seed)
        graphs[f"dense{seed}"] = (sorted_symmetric_edge_index(g), n)
    for name in ("cornell", "wisconsin"):
        graphs[name] = named_graph(name)
    pack = {"names": np.array(sorted(graphs))}
    n_gamma = 0
    for name, (ei, n) in graphs.items():
        G = graph_from_edge_index(ei, n)
        m = ei[0] < ei[1]
        edges = np.stack([ei[0][m], ei[1][m]], axis=1)
        ints = np.zeros((edges.shape[0], 6), dtype=np.int64)
        vals = np.zeros(edges.shape[0], dtype=np.float64)
        for q, (a, b) in enumerate(edges):
            val, loc = _bfc_edge_with_locals(G, int(a), int(b))
            vals[q] = float(val)
            ints[q, 0], ints[q, 1] = loc["deg1"], loc["deg2"]
            if "triangles" in loc:
                ints[q, 2:5] = len(loc["triangles"]), len(loc["squares_1"]), len(loc["squares_2"])
            if "gamma" in loc:
                ints[q, 5] = int(loc["gamma"])
                n_gamma += 1
        pack[f"{name}/edge_index"] = ei
        pack[f"{name}/n"] = np.int64(n)
        pack[f"{name}/edges"] = edges
        pack[f"{name}/ints"] = ints
        pack[f"{name}/bfc"] = vals
    np.savez_compressed(out, **pack)
    print("paper_ints_kat:", len(graphs), "graphs,", n_gamma, "edges with gamma")


def dense_from_edge_index(ei, n):
    A = torch.zeros(n, n)
    A[ei[0], ei[1]] = 1.0
    return A


def gen_cuda(out):
    graphs = {k: (sorted_symmetric_edge_index(g), g.number_of_nodes()) for k, g in toy_graphs().items()}
    for seed in range(14):
        n = 10 + 2 * seed
        g = nx.gnp_random_graph(n, 0.2 + 0.03 * (seed % 5), seed=100 + seed)
        graphs[f"gnp{seed}"] = (sorted_symmetric_edge_index(g), n)
    pack = {"names": np.array(sorted(graphs))}
    rng = np.random.default_rng(7)
    for name, (ei, n) in graphs.items():
        t0 = time.time()
        A = dense_from_edge_index(ei, n)
        C = ref_cuda.balanced_forman_curvature(A.clone())
        pack[f"{name}/edge_index"] = ei
        pack[f"{name}/n"] = np.int64(n)
        pack[f"{name}/C"] = C.numpy().copy()
        # post_delta on the argmin edge (what SDRF asks for) and on two random edges, SDRF-style neighbour lists
        G = graph_from_edge_index(ei, n)
        ix = int(C.argmin())
        picks = [(ix // n, ix % n)]
        m = np.flatnonzero(ei[0] < ei[1])
        for e in rng.choice(m, size=min(2, m.size), replace=False):
            picks.append((int(ei[0][e]), int(ei[1][e])))
        for q, (x, y) in enumerate(picks):
            xn = list(G.neighbors(x)) + [x]
            yn = list(G.neighbors(y)) + [y]
            D = ref_cuda.balanced_forman_post_delta(A.clone(), x, y, xn, yn)
            pack[f"{name}/pd{q}/xy"] = np.array([x, y], dtype=np.int64)
            pack[f"{name}/pd{q}/xn"] = np.array(xn, dtype=np.int64)
            pack[f"{name}/pd{q}/yn"] = np.array(yn, dtype=np.int64)
            pack[f"{name}/pd{q}/D"] = D.numpy().copy()
        pack[f"{name}/npd"] = np.int64(len(picks))
        print(f"  cuda {name}: n={n} {time.time() - t0:.1f}s", flush=True)
    np.savez_compressed(out, **pack)
    print("cuda_kat:", len(graphs), "graphs")


def run_reference_rewire(ei, n, loops, bound, tau, seed):
    """Unmodified ``rewire(..., 'bfc', ...)``; returns (edge_index_out, mutation log, uniforms)."""
    log = []
    orig_add, orig_rm = nx.Graph.add_edge, nx.Graph.remove_edge

    def add_edge(self, u, v, **kw):
        log.append((1, int(u), int(v)))
        return orig_add(self, u, v, **kw)

    def remove_edge(self, u, v):
        log.append((-1, int(u), int(v)))
        return orig_rm(self, u, v)

    data = Data(edge_index=torch.from_numpy(ei).long())
    data.num_nodes = n
    np.random.seed(seed)
    uniforms = np.random.RandomState(seed).random_sample(loops)
    nx.Graph.add_edge, nx.Graph.remove_edge = add_edge, remove_edge
    try:
        out = ref_rewire(data, "bfc", loops, bound, tau)
    finally:
        nx.Graph.add_edge, nx.Graph.remove_edge = orig_add, orig_rm
    return out.numpy().copy(), np.array(log, dtype=np.int64).reshape(-1, 3), uniforms


def gen_sdrf(out):
    cases = []
    # (name, graph, loops, removal_bound, tau, seed)
    cases.append(("gnp14_greedy", nx.gnp_random_graph(14, 0.25, seed=3), 6, 0.5, float("inf"), 11))
    cases.append(("gnp14_tau12", nx.gnp_random_graph(14, 0.25, seed=3), 6, 0.5, 12, 12))
    cases.append(("gnp16_tau3", nx.gnp_random_graph(16, 0.22, seed=5), 6, 0.3, 3, 13))
    cases.append(("barbell_greedy", nx.barbell_graph(5, 2), 6, 1.2, float("inf"), 14))
    cases.append(("grid_tau50", relabel(nx.grid_2d_graph(3, 4)), 5, 0.1, 50, 15))
    cases.append(("tree_greedy", nx.balanced_tree(2, 3), 5, 0.9, float("inf"), 16))
    cases.append(("gnp18_nobound", nx.gnp_random_graph(18, 0.2, seed=9), 6, 100.0, float("inf"), 17))
    pack = {"names": np.array([c[0] for c in cases])}
    for name, g, loops, bound, tau, seed in cases:
        t0 = time.time()
        ei = sorted_symmetric_edge_index(g)
        n = g.number_of_nodes()
        eo, log, uni = run_reference_rewire(ei, n, loops, bound, tau, seed)
        pack[f"{name}/edge_index"] = ei
        pack[f"{name}/n"] = np.int64(n)
        pack[f"{name}/loops"] = np.int64(loops)
        pack[f"{name}/bound"] = np.float64(bound)
        pack[f"{name}/tau"] = np.float64(tau)
        pack[f"{name}/uniforms"] = uni
        pack[f"{name}/out"] = eo
        pack[f"{name}/log"] = log
        print(f"  sdrf {name}: n={n} loops={loops} log={len(log)} {time.time() - t0:.1f}s", flush=True)
    np.savez_compressed(out, **pack)
    print("sdrf_seq:", len(cases), "cases")


def run_reference_directed(ei, n, loops, bound, tau, seed):
    """Unmodified ``sdrf_cuda_bfc(..., is_undirected=False)`` (rewiring/sdrf_cuda_bfc.py:47-49,72-73,87-88): ``G`` is a
    ``DiGraph``; the add/remove sequence is captured on ``nx.DiGraph``."""
    log = []
    orig_add, orig_rm = nx.DiGraph.add_edge, nx.DiGraph.remove_edge

    def add_edge(self, u, v, **kw):
        log.append((1, int(u), int(v)))
        return orig_add(self, u, v, **kw)

    def remove_edge(self, u, v):
        log.append((-1, int(u), int(v)))
        return orig_rm(self, u, v)

    data = Data(edge_index=torch.from_numpy(ei).long())
    data.num_nodes = n
    np.random.seed(seed)
    uniforms = np.random.RandomState(seed).random_sample(loops)
    # to_networkx builds the DiGraph with add_edge as well: patch only around the loop's own mutations by recording
    # the length of the log after set-up
    nx.DiGraph.add_edge, nx.DiGraph.remove_edge = add_edge, remove_edge
    try:
        out = ref_sdrf.sdrf_cuda_bfc(data, loops, True, bound, tau, False)
    finally:
        nx.DiGraph.add_edge, nx.DiGraph.remove_edge = orig_add, orig_rm
    # the first ei.shape[1] additions are to_networkx's (:31); from_networkx rebuilds the graph once more at the end
    # (convert_node_labels_to_integers -> add_edges_from, which does not call add_edge)
    loop_log = log[ei.shape[1]:]
    return out.edge_index.numpy().copy(), np.array(loop_log, dtype=np.int64).reshape(-1, 3), uniforms


def gen_sdrf_directed(out):
    rng = np.random.default_rng(21)
    cases = []
    for q, (n, m, loops, bound, tau, seed) in enumerate([(12, 30, 6, 0.3, 5, 31), (14, 40, 6, 0.5, float("inf"), 32),
                                                         (16, 45, 6, 0.2, 20, 33), (10, 40, 5, 0.8, float("inf"), 34),
                                                         (15, 36, 6, 100.0, 3, 35)]):
        ei = rng.integers(0, n, size=(2, m))
        ei = ei[:, ei[0] != ei[1]]
        ei = np.unique(ei, axis=1)                      # sorted, no duplicate directed edges
        ei = ei[:, rng.permutation(ei.shape[1])]        # insertion order is not sorted order
        cases.append((f"dir{q}", ei, n, loops, bound, tau, seed))
    pack = {"names": np.array([c[0] for c in cases])}
    for name, ei, n, loops, bound, tau, seed in cases:
        t0 = time.time()
        eo, log, uni = run_reference_directed(ei, n, loops, bound, tau, seed)
        pack[f"{name}/edge_index"] = ei
        pack[f"{name}/n"] = np.int64(n)
        pack[f"{name}/loops"] = np.int64(loops)
        pack[f"{name}/bound"] = np.float64(bound)
        pack[f"{name}/tau"] = np.float64(tau)
        pack[f"{name}/uniforms"] = uni
        pack[f"{name}/out"] = eo
        pack[f"{name}/log"] = log
        print(f"  sdrf-directed {name}: n={n} loops={loops} log={len(log)} {time.time() - t0:.1f}s", flush=True)
    np.savez_compressed(out, **pack)
    print("sdrf_directed_seq:", len(cases), "cases")


def run_reference_classical(ei, n, curv_type, loops, bound, tau, seed):
    """Unmodified ``rewire(data, curv_type, ...)`` -> ``sdrf_no_cuda`` (rewiring/sdrf_no_cuda.py:9-68).  The scoring loop
    probes every candidate with ``add_edge`` / ``remove_edge`` (:42-45); a probe is an addition IMMEDIATELY followed by
    the removal of the same pair and is dropped from the log, what remains are the loop's own mutations (:51, :63)."""
    log = []
    orig_add, orig_rm = nx.Graph.add_edge, nx.Graph.remove_edge

    def add_edge(self, u, v, **kw):
        log.append((1, int(u), int(v)))
        return orig_add(self, u, v, **kw)

    def remove_edge(self, u, v):
        if log and log[-1] == (1, int(u), int(v)):
            log.pop()
        else:
            log.append((-1, int(u), int(v)))
        return orig_rm(self, u, v)

    data = Data(edge_index=torch.from_numpy(ei).long())
    data.num_nodes = n
    data.x = torch.arange(n, dtype=torch.float32).view(n, 1).repeat(1, 2)     # to_networkx(..., node_attrs=['x']) (:19)
    np.random.seed(seed)
    uniforms = np.random.RandomState(seed).random_sample(loops)
    nx.Graph.add_edge, nx.Graph.remove_edge = add_edge, remove_edge
    try:
        out = ref_rewire(data, curv_type, loops, bound, tau)
    finally:
        nx.Graph.add_edge, nx.Graph.remove_edge = orig_add, orig_rm
    setup = int((ei[1] <= ei[0]).sum())            # to_networkx's own add_edge calls (columns with v <= u)
    return out.numpy().copy(), np.array(log[setup:], dtype=np.int64).reshape(-1, 3), uniforms


def gen_sdrf_classical(out):
    cases = []
    # (name, graph, curv_type, loops, removal_bound, tau, seed)
    for ct, bound in (("1d", -3.0), ("augmented", 0.5), ("haantjes", 0.5)):
        cases.append((f"gnp14_{ct}_greedy", nx.gnp_random_graph(14, 0.25, seed=3), ct, 8, bound, float("inf"), 41))
        cases.append((f"gnp16_{ct}_tau2", nx.gnp_random_graph(16, 0.22, seed=5), ct, 8, bound, 2, 42))
        cases.append((f"barbell_{ct}_tau1", nx.barbell_graph(5, 2), ct, 8, bound, 1, 43))
        cases.append((f"tree_{ct}_nobound", nx.balanced_tree(2, 3), ct, 6, 100.0, 3, 44))
    pack = {"names": np.array([c[0] for c in cases])}
    for name, g, ct, loops, bound, tau, seed in cases:
        t0 = time.time()
        ei = sorted_symmetric_edge_index(g)
        n = g.number_of_nodes()
        eo, log, uni = run_reference_classical(ei, n, ct, loops, bound, tau, seed)
        pack[f"{name}/edge_index"] = ei
        pack[f"{name}/n"] = np.int64(n)
        pack[f"{name}/curv_type"] = np.array(ct)
        pack[f"{name}/loops"] = np.int64(loops)
        pack[f"{name}/bound"] = np.float64(bound)
        pack[f"{name}/tau"] = np.float64(tau)
        pack[f"{name}/uniforms"] = uni
        pack[f"{name}/out"] = eo
        pack[f"{name}/log"] = log
        print(f"  sdrf-classical {name}: n={n} loops={loops} log={len(log)} {time.time() - t0:.1f}s", flush=True)
    np.savez_compressed(out, **pack)
    print("sdrf_classical_seq:", len(cases), "cases")


def gen_sdrf_selfloops(out):
    """Inputs WITH self-loops (the WebKB datasets carry some): the reference keeps them in ``G`` (sdrf_cuda_bfc.py:31) but
    not in ``A`` (:29), so a node with a self-loop appears twice in its own candidate list (:45-46) and the loop survives
    into the output ``edge_index`` (:93)."""
    rng = np.random.default_rng(77)
    cases = []
    for q, (g, loops, bound, tau, seed, n_self) in enumerate([
            (nx.gnp_random_graph(14, 0.25, seed=3), 8, 0.5, float("inf"), 51, 3),
            (nx.gnp_random_graph(16, 0.22, seed=5), 8, 0.3, 4, 52, 4),
            (nx.barbell_graph(5, 2), 8, 1.2, 7, 53, 2),
            (nx.gnp_random_graph(18, 0.2, seed=9), 8, 100.0, float("inf"), 54, 18)]):
        ei = sorted_symmetric_edge_index(g)
        n = g.number_of_nodes()
        who = rng.choice(n, size=min(n_self, n), replace=False)
        ei = np.concatenate([ei, np.stack([who, who])], axis=1)
        ei = ei[:, rng.permutation(ei.shape[1])]            # self-loops at arbitrary insertion positions
        cases.append((f"self{q}", ei, n, loops, bound, tau, seed))
    pack = {"names": np.array([c[0] for c in cases])}
    for name, ei, n, loops, bound, tau, seed in cases:
        eo, log, uni = run_reference_rewire(ei, n, loops, bound, tau, seed)
        pack[f"{name}/edge_index"] = ei
        pack[f"{name}/n"] = np.int64(n)
        pack[f"{name}/loops"] = np.int64(loops)
        pack[f"{name}/bound"] = np.float64(bound)
        pack[f"{name}/tau"] = np.float64(tau)
        pack[f"{name}/uniforms"] = uni
        pack[f"{name}/out"] = eo
        pack[f"{name}/log"] = log
        print(f"  sdrf-selfloops {name}: n={n} loops={loops} log={len(log)}", flush=True)
    np.savez_compressed(out, **pack)
    print("sdrf_selfloop_seq:", len(cases), "cases")


def gen_sdrf_classical_selfloops(out):
    """sdrf_no_cuda on inputs WITH self-loops: a loop is an edge of G.edges with a curvature of its own (degree counts it
    twice, the node is its own neighbour), can be the minimum / maximum edge and can be removed."""
    rng = np.random.default_rng(123)
    cases = []
    q = 0
    for ct, bound in (("1d", -3.0), ("augmented", 0.5), ("haantjes", 0.5)):
        for g, loops, tau, seed, n_self in ((nx.gnp_random_graph(14, 0.25, seed=3), 8, float("inf"), 71, 3),
                                            (nx.gnp_random_graph(16, 0.22, seed=5), 8, 2, 72, 5),
                                            (nx.barbell_graph(5, 2), 8, 1, 73, 12)):
            ei = sorted_symmetric_edge_index(g)
            n = g.number_of_nodes()
            who = rng.choice(n, size=min(n_self, n), replace=False)
            ei = np.concatenate([ei, np.stack([who, who])], axis=1)
            ei = ei[:, rng.permutation(ei.shape[1])]
            cases.append((f"cself{q}_{ct}", ei, n, ct, loops, bound, tau, seed))
            q += 1
    pack = {"names": np.array([c[0] for c in cases])}
    for name, ei, n, ct, loops, bound, tau, seed in cases:
        eo, log, uni = run_reference_classical(ei, n, ct, loops, bound, tau, seed)
        for key, val in (("edge_index", ei), ("n", np.int64(n)), ("curv_type", np.array(ct)), ("loops", np.int64(loops)),
                         ("bound", np.float64(bound)), ("tau", np.float64(tau)), ("uniforms", uni), ("out", eo), ("log", log)):
            pack[f"{name}/{key}"] = val
        print(f"  sdrf-classical-selfloops {name}: n={n} log={len(log)}", flush=True)
    np.savez_compressed(out, **pack)


def gen_sdrf_directed_selfloops(out):
    """is_undirected=False on digraphs WITH self-loops: kept in the DiGraph (successors AND predecessors of the node), not
    in A."""
    rng = np.random.default_rng(91)
    cases = []
    for q, (n, m, loops, bound, tau, seed, n_self) in enumerate([(12, 30, 6, 0.3, 5, 61, 3), (14, 40, 6, 0.5, float("inf"), 62, 4),
                                                                  (10, 36, 6, -5.0, float("inf"), 63, 10)]):
        ei = rng.integers(0, n, size=(2, m))
        ei = ei[:, ei[0] != ei[1]]
        ei = np.unique(ei, axis=1)
        who = rng.choice(n, size=n_self, replace=False)
        ei = np.concatenate([ei, np.stack([who, who])], axis=1)
        ei = ei[:, rng.permutation(ei.shape[1])]
        cases.append((f"dself{q}", ei, n, loops, bound, tau, seed))
    pack = {"names": np.array([c[0] for c in cases])}
    for name, ei, n, loops, bound, tau, seed in cases:
        eo, log, uni = run_reference_directed(ei, n, loops, bound, tau, seed)
        for key, val in (("edge_index", ei), ("n", np.int64(n)), ("loops", np.int64(loops)), ("bound", np.float64(bound)),
                         ("tau", np.float64(tau)), ("uniforms", uni), ("out", eo), ("log", log)):
            pack[f"{name}/{key}"] = val
        print(f"  sdrf-directed-selfloops {name}: n={n} loops={loops} log={len(log)}", flush=True)
    np.savez_compressed(out, **pack)


if __name__ == "__main__":
    JOBS = [("sdrf_classical_selfloops", gen_sdrf_classical_selfloops, "sdrf_classical_selfloop_seq.npz"),
            ("sdrf_directed_selfloops", gen_sdrf_directed_selfloops, "sdrf_directed_selfloop_seq.npz"),
            ("sdrf_selfloops", gen_sdrf_selfloops, "sdrf_selfloop_seq.npz"),
            ("sdrf_classical", gen_sdrf_classical, "sdrf_classical_seq.npz"),
            ("paper_ints", gen_paper_ints, "paper_ints_kat.npz"),
            ("sdrf_directed", gen_sdrf_directed, "sdrf_directed_seq.npz"),
            ("paper", gen_paper, "paper_kat.npz"),
            ("cuda", gen_cuda, "cuda_kat.npz"),
            ("sdrf", gen_sdrf, "sdrf_seq.npz")]
    which = sys.argv[1:] or [name for name, _, _ in JOBS]
    unknown = set(which) - {name for name, _, _ in JOBS}
    if unknown:
        raise SystemExit(f"unknown fixture(s) {sorted(unknown)}; choose from {[name for name, _, _ in JOBS]}")
    for name, fn, out in JOBS:
        if name in which:
            fn(os.path.join(HERE, out))

"""CPU suite, part 2: host-side logic of the package (graph set-up, drop-in softmax, C-ABI surface)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import PKG, REPO
from dcr import graph as G
from dcr.synth import SHAPES, chung_lu_graph, csr_from_edge_index
from oracle.sdrf import dense_adjacency, edge_index_from_adjacency, networkx_adjacency


def test_graph_setup_matches_networkx_semantics():
    rng = np.random.default_rng(0)
    for trial in range(150):
        n = int(rng.integers(2, 30))
        ei = rng.integers(0, n, size=(2, int(rng.integers(1, 80))))
        ei = ei[:, ei[0] != ei[1]]
        if ei.shape[1] == 0:
            continue
        if trial % 3 == 0:   # sorted + symmetric, the form the reference's datasets come in
            ei = np.concatenate([ei, ei[::-1]], axis=1)
            k = np.unique(ei[0] * n + ei[1])
            ei = np.stack([k // n, k % n])
        adj = networkx_adjacency(ei, n)
        rp, od = G.networkx_order(ei, n)
        assert [list(a) for a in adj] == [od[rp[v]:rp[v + 1]].tolist() for v in range(n)]
        assert np.array_equal(G.from_networkx_order(rp, od), edge_index_from_adjacency(adj))
        rp2, c2 = G.undirected_csr(ei)
        A = dense_adjacency(ei)
        for v in range(A.shape[0]):
            assert np.array_equal(np.flatnonzero(A[v]), c2[rp2[v]:rp2[v + 1]])


def test_networkx_order_against_real_networkx():
    nx = pytest.importorskip("networkx")
    rng = np.random.default_rng(1)
    for trial in range(30):
        n = int(rng.integers(3, 25))
        ei = rng.integers(0, n, size=(2, int(rng.integers(2, 60))))
        ei = ei[:, ei[0] != ei[1]]
        D = nx.DiGraph()
        D.add_nodes_from(range(n))
        for u, v in ei.T.tolist():
            D.add_edge(u, v)
        U = D.to_undirected()
        rp, od = G.networkx_order(ei, n)
        assert [list(U.neighbors(v)) for v in range(n)] == [od[rp[v]:rp[v + 1]].tolist() for v in range(n)]


def test_sorted_symmetric_input_gives_ascending_adjacency():
    ei = chung_lu_graph(60, 150, 0.8, 0.2, 3)
    rp, od = G.networkx_order(ei, 60)
    for v in range(60):
        row = od[rp[v]:rp[v + 1]]
        assert np.all(np.diff(row) > 0)


def test_synthetic_graphs_have_the_named_shapes():
    for name in ("cornell", "texas", "wisconsin", "cora"):
        n, e, alpha, p_tri, seed = SHAPES[name]
        ei = chung_lu_graph(n, e, alpha, p_tri, seed)
        assert ei.shape == (2, 2 * e)
        assert int(ei.max()) < n
        key = ei[0] * n + ei[1]
        assert np.all(np.diff(key) > 0)                       # sorted, no duplicates
        assert np.array_equal(np.sort(ei[1] * n + ei[0]), key)  # symmetric
        assert not np.any(ei[0] == ei[1])
        rowptr, col = csr_from_edge_index(ei, n)
        assert (np.diff(rowptr) > 0).all()                    # no isolated nodes
        assert np.array_equal(chung_lu_graph(n, e, alpha, p_tri, seed), ei)   # deterministic


def test_dropin_softmax_matches_reference_semantics():
    from oracle.sdrf import softmax as oracle_softmax
    from utils.softmax import softmax
    rng = np.random.default_rng(3)
    for tau in (float("inf"), 1, 12, 145):
        a = rng.normal(size=17)
        assert np.array_equal(softmax(a, tau), oracle_softmax(a, tau))
    a = np.array([0.5, 0.5, 0.1])
    assert softmax(a, float("inf")).tolist() == [1.0, 0.0, 0.0]    # first argmax


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "dcr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dcr_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    from dcr import lib as L
    path = L.LIB_PATH
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(path)
    names = _declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/dcr.h but not exported"
    assert set(L.SIGNATURES) == set(names)
    assert L.load().dcr_version() == 1


def test_product_path_never_imports_the_oracle():
    # the oracle is test infrastructure: nothing under the package may reference it
    for root, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(root, f)


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from dcr import lib as L
    from dcr.bfc import DeviceCSR
    with pytest.raises(L.DcrError):
        DeviceCSR.from_host(np.array([0, 1, 2], dtype=np.int32), np.array([1, 0], dtype=np.int32))


def test_dense_entry_points_fail_loudly_without_cuda():
    """No CPU fallback anywhere on the dense signatures: small-graph path, CSR route and post_delta all raise."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without CUDA")
    from curvature.bfc_cuda import balanced_forman_curvature, balanced_forman_post_delta
    from dcr import bfc
    from dcr.lib import DcrError
    A = torch.zeros(6, 6)
    A[0, 1] = A[1, 0] = 1
    with pytest.raises(DcrError):
        balanced_forman_curvature(A)
    with pytest.raises(DcrError):
        bfc.cuda_flavour_dense_small(A, torch.zeros(6, 6))
    with pytest.raises(DcrError):
        balanced_forman_post_delta(A, 0, 1, [1, 0], [0, 1])
    big = torch.zeros(1100, 1100)                       # above the small-graph path: the CSR route
    with pytest.raises(DcrError):
        balanced_forman_curvature(big)


def test_directed_and_classical_graph_setup_match_the_oracle_builders():
    """dcr.graph.digraph_order / from_digraph_order / classical_order (vectorised numpy, product host code) against the
    dict-based restatements of networkx / PyG behaviour in the oracle, on shuffled inputs."""
    import warnings
    from dcr import graph as G
    from oracle.sdrf import edge_index_from_digraph, networkx_digraph
    from oracle.sdrf_classical import networkx_adjacency_classical
    from helpers import gnp
    rng = np.random.default_rng(0)
    for t in range(12):
        n = 15 + t
        ei = rng.integers(0, n, size=(2, 5 * n))
        ei = ei[:, ei[0] != ei[1]]
        ei = np.unique(ei, axis=1)
        ei = ei[:, rng.permutation(ei.shape[1])]
        s_rp, s_ord, p_rp, p_ord = G.digraph_order(ei, n)
        succ, pred = networkx_digraph(ei, n)
        for v in range(n):
            assert list(succ[v]) == s_ord[s_rp[v]:s_rp[v + 1]].tolist()
            assert list(pred[v]) == p_ord[p_rp[v]:p_rp[v + 1]].tolist()
        assert np.array_equal(G.from_digraph_order(s_rp, s_ord), edge_index_from_digraph(succ))
        und = gnp(n, 0.3, t)
        und = und[:, rng.permutation(und.shape[1])]
        rp, od = G.classical_order(und, n)
        adj = networkx_adjacency_classical(und, n)
        for v in range(n):
            assert list(adj[v]) == od[rp[v]:rp[v + 1]].tolist()
    # only the columns with v <= u survive to_networkx(..., to_undirected=True): an upper-triangle-only input is empty
    up = gnp(20, 0.3, 1)
    up = up[:, up[0] < up[1]]
    rp, od = G.classical_order(up, 20)
    assert od.size == 0 and rp[-1] == 0
    with pytest.raises(NotImplementedError):           # a repeated directed edge would make A weighted
        G.digraph_order(np.array([[0, 0, 1], [1, 1, 2]]), 3)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        G.digraph_order(np.array([[0, 1, 1], [1, 1, 2]]), 3)
        assert any("self-loops" in str(x.message) for x in w)


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs first): one JSON line with the contract keys, the identical
    `config` dict our arm uses, and no libdcr.so in the process."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--workload", "cora",
                        "--steps", "1", "--warmup", "0", "--cpu-budget", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "bfc_edges_per_sec" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference")
    sys.path.insert(0, REPO)
    import bench
    assert d["config"] == bench.base_config("cora", d["config"]["nodes"], d["config"]["undirected_edges"], 1)

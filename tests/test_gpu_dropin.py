"""GPU: the drop-in modules keep the reference's module paths, signatures and return conventions."""
import numpy as np
import pytest

from helpers import dense_of, gnp

pytestmark = pytest.mark.gpu


def test_balanced_forman_curvature_signature_and_inplace_semantics():
    import torch
    from curvature.bfc_cuda import balanced_forman_curvature
    from oracle.cuda_flavour import bfc_cuda_dense
    n = 60
    ei = gnp(n, 0.12, 1)
    A = torch.from_numpy(dense_of(ei, n)).cuda()
    C = balanced_forman_curvature(A)
    assert C.shape == (n, n) and C.dtype == torch.float32 and C.is_cuda
    ref = bfc_cuda_dense(dense_of(ei, n))["C"]
    assert np.array_equal(C.cpu().numpy().view(np.uint32), ref.view(np.uint32))
    C2 = torch.full((n, n), 3.0, device="cuda")
    out = balanced_forman_curvature(A, C=C2)
    assert out is C2 and torch.equal(C2, C)          # written in place and returned (bfc_cuda.py:56-57,65)


def test_small_dense_path_bitwise_and_validation():
    """n <= 1024: bit-packed one-kernel path (dcr_dense_small.cu) against the compiled-dataflow oracle; inputs outside
    the covered domain raise like the CSR route does."""
    import torch
    from curvature.bfc_cuda import balanced_forman_curvature
    from dcr import bfc
    from helpers import toy_graphs
    from oracle.cuda_flavour import bfc_cuda_dense
    graphs = dict(toy_graphs())
    for s, (n, p) in enumerate([(33, 0.3), (64, 0.1), (257, 0.05), (700, 0.02), (1024, 0.01)]):
        graphs[f"gnp{n}"] = (gnp(n, p, 40 + s), n)
    for name, (ei, n) in graphs.items():
        assert n <= bfc.SMALL_DENSE_MAX_N
        An = dense_of(ei, n)
        C = torch.full((n, n), 7.0, device="cuda")
        out = balanced_forman_curvature(torch.from_numpy(An).cuda(), C=C)
        assert out is C
        ref = bfc_cuda_dense(An)["C"]
        assert np.array_equal(C.cpu().numpy().view(np.uint32), ref.view(np.uint32)), name
    asym = torch.zeros(8, 8, device="cuda")
    asym[0, 1] = asym[1, 2] = asym[2, 0] = asym[0, 2] = 1   # asymmetric: the directed route (tests/test_gpu_directed.py)
    got = balanced_forman_curvature(asym).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), bfc_cuda_dense(asym.cpu().numpy())["C"].view(np.uint32))
    bad = torch.zeros(8, 8, device="cuda")
    bad[2, 2] = 1                                        # self-loop
    keep = torch.full((8, 8), 5.0, device="cuda")
    with pytest.raises(NotImplementedError):
        balanced_forman_curvature(bad, C=keep)
    assert bool((keep == 5.0).all())                     # a rejected input leaves the caller's C untouched
    bad = torch.zeros(8, 8, device="cuda")
    bad[0, 1] = bad[1, 0] = 2                            # weighted
    with pytest.raises(NotImplementedError):
        balanced_forman_curvature(bad)


def test_balanced_forman_post_delta_signature():
    import torch
    from curvature.bfc_cuda import balanced_forman_post_delta
    from oracle.cuda_flavour import post_delta_dense
    n = 30
    ei = gnp(n, 0.2, 2)
    An = dense_of(ei, n)
    A = torch.from_numpy(An).cuda()
    x, y = int(ei[0][0]), int(ei[1][0])
    xn = np.flatnonzero(An[x]).tolist() + [x]
    yn = np.flatnonzero(An[y]).tolist() + [y]
    D = balanced_forman_post_delta(A, x, y, xn, yn)
    assert D.shape == (len(xn), len(yn))
    ref = post_delta_dense(An, x, y, xn, yn)
    assert np.array_equal(D.cpu().numpy().view(np.uint32), ref.view(np.uint32))


def test_bfc_naive_dropin_on_networkx_graph():
    nx = pytest.importorskip("networkx")
    from curvature.bfc_naive import bfc, bfc_edge
    from oracle.paper_flavour import adjacency_sets, bfc_edge_fields
    G = nx.gnp_random_graph(40, 0.15, seed=4)
    ei = np.array([(u, v) for u, v in G.edges] + [(v, u) for u, v in G.edges]).T
    adj = adjacency_sets(ei, 40)
    bfc(G)
    for u, v in G.edges:
        want = bfc_edge_fields(adj, u, v)[6]
        assert G[u][v]["bfc"] == want
        assert type(G[u][v]["bfc"]) is type(want)     # int 0 when deg_min == 1, else float
    u, v = next(iter(G.edges))
    assert bfc_edge(G, u, v) == bfc_edge_fields(adj, u, v)[6]
    from curvature.bfc_naive import prepare
    P = prepare(G)                                       # repeated queries on an unchanged graph: CSR uploaded once
    for a, b in list(G.edges)[:10]:
        assert bfc_edge(P, a, b) == bfc_edge_fields(adj, a, b)[6]


def test_rewire_dropin_returns_edge_index_and_consumes_numpy_stream():
    import torch
    from dcr.synth import named_graph
    from oracle.sdrf import sdrf_oracle
    from rewiring.rewire import rewire
    from torch_geometric.data import Data
    ei, n = named_graph("texas")
    data = Data(edge_index=torch.from_numpy(ei))
    data.num_nodes = n
    loops = 20
    np.random.seed(123)
    out = rewire(data, "bfc", loops, 1.64, 22)
    after = np.random.random_sample()
    assert isinstance(out, torch.Tensor) and out.dtype == torch.long and out.shape[0] == 2
    uni = np.random.RandomState(123).random_sample(loops + 1)
    want, wlog = sdrf_oracle(ei, n, loops, True, 1.64, 22, uni)
    assert np.array_equal(out.numpy(), want)
    draws = sum(r["n_candidates"] > 0 for r in wlog)
    assert after == uni[draws]        # the global generator advanced by exactly the reference's number of draws
    assert rewire(data, None, loops, 1.64, 22) is data.edge_index


def test_balanced_forman_curvature_dense_regime_uses_tensor_path_and_matches_oracle():
    import torch
    from curvature import bfc_cuda
    from dcr.synth import chung_lu_graph
    from oracle.cuda_flavour import bfc_cuda_dense
    n = 1100                                            # above the small-graph path (n <= 1024)
    ei = chung_lu_graph(n, 30000, 0.5, 0.3, 4)         # average degree 55: dense regime
    An = dense_of(ei, n)
    from dcr import bfc
    assert bfc_cuda._dense_regime(bfc.DeviceCSR.from_dense(torch.from_numpy(An).cuda()))
    C = bfc_cuda.balanced_forman_curvature(torch.from_numpy(An).cuda())
    ref = bfc_cuda_dense(An)["C"]
    assert np.array_equal(C.cpu().numpy().view(np.uint32), ref.view(np.uint32))


def test_config2_rewire_then_gcn_on_gpu():
    """Config 2: SDRF (greedy and softmax tau) on the wisconsin-shaped graph, then a GCN trained on the B200 on the
    rewired edge_index (GCNConv stand-in; the reference's models/gcn.py runs unchanged on it — CPU test)."""
    import torch
    from dcr import compat
    from dcr.synth import SDRF_PARAMS, named_graph
    from rewiring.rewire import rewire
    compat.ensure_torch_geometric()
    from torch_geometric.data import Data
    from torch_geometric.nn import GCNConv
    ei, n = named_graph("wisconsin")
    loops, tau, bound = SDRF_PARAMS["wisconsin"]
    for t in (tau, float("inf")):
        data = Data(edge_index=torch.from_numpy(ei))
        data.num_nodes = n
        np.random.seed(5)
        new_ei = rewire(data, "bfc", loops, bound, t).cuda()
        assert new_ei.shape[0] == 2 and int(new_ei.max()) < n
        g = torch.Generator().manual_seed(0)
        y = torch.randint(0, 3, (n,), generator=g)
        x = (torch.randn(n, 16, generator=g)
             + 2.0 * torch.nn.functional.one_hot(y, 3).float() @ torch.randn(3, 16, generator=g)).cuda()
        y = y.cuda()
        torch.manual_seed(0)
        c1, c2 = GCNConv(16, 32).cuda(), GCNConv(32, 3).cuda()
        opt = torch.optim.Adam(list(c1.parameters()) + list(c2.parameters()), lr=0.05)
        first = None
        for _ in range(60):
            opt.zero_grad()
            logits = c2(torch.relu(c1(x, new_ei)), new_ei)
            loss = torch.nn.functional.cross_entropy(logits, y)
            loss.backward()
            opt.step()
            first = loss.item() if first is None else first
        assert loss.item() < 0.5 * first


def test_post_delta_large_candidate_matrix_uses_several_ctas():
    """(deg x + 1)(deg y + 1) > 4096 cells: the cells kernel runs on several CTAs (one CTA prepares the shared terms);
    symmetric and asymmetric A against the dense oracle, bit for bit."""
    import torch
    from curvature.bfc_cuda import balanced_forman_post_delta
    from oracle.cuda_flavour import post_delta_dense
    n = 420
    ei = gnp(n, 0.3, 77)
    An = dense_of(ei, n)
    rng = np.random.default_rng(1)
    for directed in (False, True):
        if directed:
            An = An * (rng.random((n, n)) < 0.7)          # drop 30 % of the entries: asymmetric
            np.fill_diagonal(An, 0)
        A = torch.from_numpy(An.astype(np.float32)).cuda()
        nz = np.argwhere(An > 0)
        for x, y in [tuple(nz[len(nz) // 3]), tuple(nz[-1]), (0, 0)]:
            xn = np.flatnonzero(An[x]).tolist() + [int(x)]
            yn = (np.flatnonzero(An[:, y]) if directed else np.flatnonzero(An[y])).tolist() + [int(y)]
            assert len(xn) * len(yn) > 4096
            want = post_delta_dense(An.astype(np.float32), int(x), int(y), xn, yn, "compiled")
            got = balanced_forman_post_delta(A, int(x), int(y), xn, yn).cpu().numpy()
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (directed, x, y)

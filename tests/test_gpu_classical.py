"""GPU parity, classical-curvature SDRF (rewiring/sdrf_no_cuda.py:9-68): the LOOP_CLASSICAL flavour of
the device loop against the goldens of the UNMODIFIED reference and the oracle restatement."""
import numpy as np
import pytest

from helpers import gnp, golden, toy_graphs

pytestmark = pytest.mark.gpu


def _oracle_log_tuples(log):
    return [(r["x"], r["y"], r["n_candidates"], r["k"], r["l"], r["choice"],
             -1 if r["removed"] is None else r["removed"][0], -1 if r["removed"] is None else r["removed"][1])
            for r in log]


def _run_both(ei, n, ct, loops, bound, tau, seed, remove_edges=True, guard=1e-9):
    from dcr import sdrf
    from oracle.sdrf_classical import sdrf_classical_oracle
    uni = np.random.RandomState(seed).random_sample(loops)
    got, glog = sdrf.sdrf(ei, n, loops, remove_edges, bound, tau, uniforms=uni, return_log=True, curv_type=ct,
                          guard=guard)
    want, wlog = sdrf_classical_oracle(ei, n, ct, loops, remove_edges, bound, tau, uni)
    assert [tuple(int(v) for v in r) for r in glog] == _oracle_log_tuples(wlog)
    assert np.array_equal(got, want)
    return glog


def test_classical_golden_sequences_of_the_unmodified_reference():
    from dcr import sdrf
    z = golden("sdrf_classical_seq.npz")
    for name in (str(s) for s in z["names"]):
        ei, n = z[f"{name}/edge_index"], int(z[f"{name}/n"])
        got, log = sdrf.sdrf(ei, n, int(z[f"{name}/loops"]), True, float(z[f"{name}/bound"]), float(z[f"{name}/tau"]),
                             uniforms=z[f"{name}/uniforms"], return_log=True, curv_type=str(z[f"{name}/curv_type"]))
        seq = []
        for r in log:
            if r[3] >= 0:
                seq.append((1, int(r[3]), int(r[4])))
            if r[6] >= 0:
                seq.append((-1, int(r[6]), int(r[7])))
        assert np.array_equal(np.array(seq, dtype=np.int64).reshape(-1, 3), z[f"{name}/log"]), name
        assert np.array_equal(got, z[f"{name}/out"]), name


@pytest.mark.parametrize("ct,bound", [("1d", -4.0), ("augmented", 0.5), ("haantjes", 1.5)])
@pytest.mark.parametrize("tau", [float("inf"), 1, 3])
def test_classical_small_graphs_match_oracle(ct, bound, tau):
    for seed in range(3):
        n = 15 + 6 * seed
        _run_both(gnp(n, 0.2, 500 + seed), n, ct, 15, bound, tau, 70 + seed)
    for name, (ei, n) in toy_graphs().items():          # symmetric graphs: every min / max is a tie
        _run_both(ei, n, ct, 6, bound, tau, 3)


def test_classical_named_shapes_and_exits():
    from dcr.synth import named_graph
    ei, n = named_graph("cornell")
    for ct in ("1d", "augmented", "haantjes"):
        _run_both(ei, n, ct, 40, 0.5, 2, 8)
        _run_both(ei, n, ct, 30, 0.5, float("inf"), 8, remove_edges=False)
    ei, n = toy_graphs()["k5"]                          # complete graph: no candidates -> both exits
    _run_both(ei, n, "augmented", 5, 100.0, float("inf"), 1)
    _run_both(ei, n, "augmented", 5, 100.0, float("inf"), 1, remove_edges=False)
    _run_both(ei, n, "augmented", 5, 0.0, float("inf"), 1)     # removal only


def test_classical_cora_shape_long_run_and_host_redecision():
    from dcr.synth import named_graph
    ei, n = named_graph("cora")
    _run_both(ei, n, "augmented", 150, 0.5, 2, 12)
    ei, n = named_graph("texas")
    _run_both(ei, n, "1d", 30, -3.0, 1, 5, guard=2.0)          # every draw re-decided by the host with numpy


def test_classical_dropin_signature_errors_and_one_direction_quirk():
    import torch
    from rewiring.rewire import rewire
    from rewiring.sdrf_no_cuda import sdrf_no_cuda
    from torch_geometric.data import Data
    from oracle.sdrf_classical import sdrf_classical_oracle
    n = 22
    ei = gnp(n, 0.2, 17)
    data = Data(edge_index=torch.from_numpy(ei).long())
    data.num_nodes = n
    data.x = torch.arange(n, dtype=torch.float32).view(n, 1)
    np.random.seed(5)
    out = rewire(data, "haantjes", 8, 0.5, 2)
    uni = np.random.RandomState(5).random_sample(8)
    want, _ = sdrf_classical_oracle(ei, n, "haantjes", 8, True, 0.5, 2, uni)
    assert out.dtype == torch.long and np.array_equal(out.numpy(), want)
    res = sdrf_no_cuda(data, "1d", 4, True, -2.0, float("inf"), uniforms=uni)
    assert res.x is data.x and res.num_nodes == n
    # exp overflow: tau * improvement = 436 * 2 -> inf / inf = NaN -> numpy's ValueError (the reference's behaviour)
    with pytest.raises(ValueError):
        sdrf_no_cuda(data, "augmented", 4, True, 0.5, 436, uniforms=uni)
    with pytest.raises(Exception):
        sdrf_no_cuda(data, "ollivier", 4, True, 0.5, 2, uniforms=uni)
    # to_networkx(..., to_undirected=True) keeps only the columns with v <= u: an upper-triangle-only input has no edges
    upper = ei[:, ei[0] < ei[1]]
    data2 = Data(edge_index=torch.from_numpy(upper).long())
    data2.num_nodes = n
    data2.x = data.x
    with pytest.raises(ValueError):
        sdrf_no_cuda(data2, "1d", 2, True, 0.0, 1, uniforms=uni)


def test_classical_inputs_with_self_loops():
    """sdrf_no_cuda keeps self-loops: a loop is an edge of G.edges with its own curvature (it can be the minimum edge and
    it can be removed), counts twice in the degree and makes the node its own neighbour.  Goldens of the unmodified
    reference (tests/golden/sdrf_classical_selfloop_seq.npz) and the oracle on random graphs."""
    from dcr import sdrf
    z = golden("sdrf_classical_selfloop_seq.npz")
    for name in (str(s) for s in z["names"]):
        ei, n = z[f"{name}/edge_index"], int(z[f"{name}/n"])
        got, log = sdrf.sdrf(ei, n, int(z[f"{name}/loops"]), True, float(z[f"{name}/bound"]), float(z[f"{name}/tau"]),
                             uniforms=z[f"{name}/uniforms"], return_log=True, curv_type=str(z[f"{name}/curv_type"]))
        seq = []
        for r in log:
            if r[3] >= 0:
                seq.append((1, int(r[3]), int(r[4])))
            if r[6] >= 0:
                seq.append((-1, int(r[6]), int(r[7])))
        assert np.array_equal(np.array(seq, dtype=np.int64).reshape(-1, 3), z[f"{name}/log"]), name
        assert np.array_equal(got, z[f"{name}/out"]), name
    rng = np.random.default_rng(8)
    for s, (ct, bound) in enumerate([("1d", -4.0), ("augmented", 0.5), ("haantjes", 1.5), ("augmented", 100.0)]):
        n = 18 + 5 * s
        ei = gnp(n, 0.2, 700 + s)
        who = rng.choice(n, size=5 + 3 * s, replace=False)
        ei = np.concatenate([ei, np.stack([who, who])], axis=1)
        ei = ei[:, rng.permutation(ei.shape[1])]
        for tau in (float("inf"), 2):
            _run_both(ei, n, ct, 16, bound, tau, 60 + s)

"""CPU suite, part 1: the oracle against the golden outputs of the UNMODIFIED reference and SURVEY.md App. G."""
import numpy as np
import pytest

from helpers import dense_of, gnp, golden, toy_graphs
from oracle.cuda_flavour import bfc_cuda_dense, closing_value, post_delta_dense
from oracle.paper_flavour import bfc_paper
from oracle.sdrf import choice_index, sdrf_oracle, softmax


def _names(z):
    return [str(s) for s in z["names"]]


def test_paper_oracle_matches_reference_bit_for_bit():
    z = golden("paper_kat.npz")
    total = 0
    for name in _names(z):
        r = bfc_paper(z[f"{name}/edge_index"], int(z[f"{name}/n"]), z[f"{name}/edges"])
        assert np.array_equal(r["bfc"], z[f"{name}/bfc"]), name     # fp64, same operation order -> identical bits
        total += len(r["bfc"])
    assert total > 2000


def test_cuda_oracle_sim32_matches_reference_simulator_bit_for_bit():
    z = golden("cuda_kat.npz")
    for name in _names(z):
        ei, n = z[f"{name}/edge_index"], int(z[f"{name}/n"])
        A = dense_of(ei, n)
        r = bfc_cuda_dense(A, "sim32")
        assert np.array_equal(r["C"].view(np.uint32), z[f"{name}/C"].view(np.uint32)), name
        # the compiled dataflow (fp64, two roundings) stays within fp32 noise of the simulator
        rc = bfc_cuda_dense(A, "compiled")
        assert np.abs(rc["C"] - z[f"{name}/C"]).max() <= 1e-6
        for q in range(int(z[f"{name}/npd"])):
            x, y = (int(v) for v in z[f"{name}/pd{q}/xy"])
            D = post_delta_dense(A, x, y, z[f"{name}/pd{q}/xn"].tolist(), z[f"{name}/pd{q}/yn"].tolist(), "sim32")
            assert np.array_equal(D.view(np.uint32), z[f"{name}/pd{q}/D"].view(np.uint32)), (name, q)


def test_sdrf_oracle_sim32_reproduces_reference_sequences():
    z = golden("sdrf_seq.npz")
    for name in _names(z):
        ei, n = z[f"{name}/edge_index"], int(z[f"{name}/n"])
        out, log = sdrf_oracle(ei, n, int(z[f"{name}/loops"]), True, float(z[f"{name}/bound"]),
                               float(z[f"{name}/tau"]), z[f"{name}/uniforms"], rounding="sim32", verify_a2_every=1)
        mine = []
        for r in log:
            if r["k"] >= 0:
                mine.append((1, r["k"], r["l"]))
            if r["removed"] is not None:
                mine.append((-1,) + tuple(r["removed"]))
        assert np.array_equal(np.array(mine).reshape(-1, 3), z[f"{name}/log"]), name
        assert np.array_equal(out, z[f"{name}/out"]), name


def test_sdrf_oracle_directed_mode_reproduces_reference_sequences():
    """is_undirected=False (rewiring/sdrf_cuda_bfc.py:47-49,72-73,87-88): add/remove sequence and the
    output edge_index of the UNMODIFIED reference on random directed graphs with unsorted insertion order."""
    z = golden("sdrf_directed_seq.npz")
    for name in _names(z):
        ei, n = z[f"{name}/edge_index"], int(z[f"{name}/n"])
        out, log = sdrf_oracle(ei, n, int(z[f"{name}/loops"]), True, float(z[f"{name}/bound"]),
                               float(z[f"{name}/tau"]), z[f"{name}/uniforms"], rounding="sim32", verify_a2_every=1,
                               is_undirected=False)
        mine = []
        for r in log:
            if r["k"] >= 0:
                mine.append((1, r["k"], r["l"]))
            if r["removed"] is not None:
                mine.append((-1,) + tuple(r["removed"]))
        assert np.array_equal(np.array(mine).reshape(-1, 3), z[f"{name}/log"]), name
        assert np.array_equal(out, z[f"{name}/out"]), name


def _mutations(log):
    mine = []
    for r in log:
        if r["k"] >= 0:
            mine.append((1, r["k"], r["l"]))
        if r["removed"] is not None:
            mine.append((-1,) + tuple(r["removed"]))
    return np.array(mine, dtype=np.int64).reshape(-1, 3)


def test_sdrf_oracle_reproduces_reference_on_inputs_with_self_loops():
    """Quirk D.1: G keeps self-loops (sdrf_cuda_bfc.py:31), A drops them (:29) — the unmodified reference on four graphs with
    self-loops at arbitrary insertion positions (sequence + output, which still contains the loops)."""
    z = golden("sdrf_selfloop_seq.npz")
    for name in _names(z):
        ei, n = z[f"{name}/edge_index"], int(z[f"{name}/n"])
        out, log = sdrf_oracle(ei, n, int(z[f"{name}/loops"]), True, float(z[f"{name}/bound"]),
                               float(z[f"{name}/tau"]), z[f"{name}/uniforms"], rounding="sim32", verify_a2_every=1)
        assert np.array_equal(_mutations(log), z[f"{name}/log"]), name
        assert np.array_equal(out, z[f"{name}/out"]), name
        assert int((out[0] == out[1]).sum()) > 0


def test_sdrf_oracle_directed_mode_with_self_loops():
    z = golden("sdrf_directed_selfloop_seq.npz")
    for name in _names(z):
        ei, n = z[f"{name}/edge_index"], int(z[f"{name}/n"])
        out, log = sdrf_oracle(ei, n, int(z[f"{name}/loops"]), True, float(z[f"{name}/bound"]), float(z[f"{name}/tau"]),
                               z[f"{name}/uniforms"], rounding="sim32", verify_a2_every=1, is_undirected=False)
        assert np.array_equal(_mutations(log), z[f"{name}/log"]), name
        assert np.array_equal(out, z[f"{name}/out"]), name


def test_sdrf_classical_oracle_reproduces_reference_sequences():
    """rewiring/sdrf_no_cuda.py:9-68 with '1d' / 'augmented' / 'haantjes' (SURVEY.md §8f-4): add/remove sequence and
    output edge_index of the UNMODIFIED reference, greedy and stochastic."""
    from oracle.sdrf_classical import sdrf_classical_oracle
    z = golden("sdrf_classical_seq.npz")
    seen = set()
    for name in _names(z):
        ei, n = z[f"{name}/edge_index"], int(z[f"{name}/n"])
        ct = str(z[f"{name}/curv_type"])
        seen.add(ct)
        out, log = sdrf_classical_oracle(ei, n, ct, int(z[f"{name}/loops"]), True, float(z[f"{name}/bound"]),
                                         float(z[f"{name}/tau"]), z[f"{name}/uniforms"])
        assert np.array_equal(_mutations(log), z[f"{name}/log"]), name
        assert np.array_equal(out, z[f"{name}/out"]), name
    assert seen == {"1d", "augmented", "haantjes"}
    z = golden("sdrf_classical_selfloop_seq.npz")       # inputs with self-loops: loops as minimum edges, loops removed
    loop_min = loop_removed = 0
    for name in _names(z):
        ei, n = z[f"{name}/edge_index"], int(z[f"{name}/n"])
        out, log = sdrf_classical_oracle(ei, n, str(z[f"{name}/curv_type"]), int(z[f"{name}/loops"]), True,
                                         float(z[f"{name}/bound"]), float(z[f"{name}/tau"]), z[f"{name}/uniforms"])
        assert np.array_equal(_mutations(log), z[f"{name}/log"]), name
        assert np.array_equal(out, z[f"{name}/out"]), name
        loop_min += sum(r["x"] == r["y"] for r in log)
        loop_removed += sum(r["removed"] is not None and r["removed"][0] == r["removed"][1] for r in log)
    assert loop_min > 0 and loop_removed > 0


# (graph, edge) -> cuda (d_i, d_j, A2ij, sharp, lam, C) ; paper (tri, sq1, sq2, gamma, bfc)
APP_G = [
    ("path4", (0, 1), (1, 2, 0, 3, 2, 1.75), (0, 0, 0, 0, 0.0)),
    ("path4", (1, 2), (2, 2, 0, 4, 2, 1.0), (0, 0, 0, 0, 0.0)),
    ("star5", (0, 1), (4, 1, 0, 5, 4, 0.8125), (0, 0, 0, 0, 0.0)),
    ("c3", (0, 1), (2, 2, 1, 2, 2, 2.0), (1, 0, 0, 0, 1.5)),
    ("c4", (0, 1), (2, 2, 0, 4, 2, 1.0), (0, 1, 1, 1, 1.0)),
    ("c5", (0, 1), (2, 2, 0, 4, 2, 1.0), (0, 0, 0, 0, 0.0)),
    ("k4", (0, 1), (3, 3, 2, 6, 3, 2.0), (2, 0, 0, 0, 4 / 3)),
    ("k5", (0, 1), (4, 4, 3, 8, 4, 1.75), (3, 0, 0, 0, 1.25)),
    ("k33", (0, 3), (3, 3, 0, 6, 3, 0.0), (0, 2, 2, 2, 0.0)),
    ("grid3", (0, 1), (2, 3, 0, 5, 3, 2 / 9), (0, 1, 1, 1, 1 / 3)),
    ("grid3", (1, 4), (3, 4, 0, 7, 4, -0.395833), (0, 2, 2, 1, 1 / 6)),
    ("petersen", (0, 1), (3, 3, 0, 6, 3, 0.0), (0, 0, 0, 0, -2 / 3)),
    ("cube", (0, 1), (3, 3, 0, 6, 3, 0.0), (0, 2, 2, 1, 2 / 3)),
    ("barbell41", (0, 1), (3, 3, 2, 6, 3, 2.0), (2, 0, 0, 0, 4 / 3)),
    ("barbell41", (0, 3), (3, 4, 2, 7, 4, 1.270833), (2, 0, 0, 0, 5 / 6)),
    ("barbell41", (3, 4), (4, 2, 0, 6, 4, -0.125), (0, 0, 0, 0, -0.5)),
]


@pytest.mark.parametrize("gname,edge,cuda,paper", APP_G)
def test_appendix_g_known_answers(gname, edge, cuda, paper):
    ei, n = toy_graphs()[gname]
    i, j = edge
    r = bfc_cuda_dense(dense_of(ei, n), "compiled")
    di, dj, a2, sharp, lam, c = cuda
    assert (int(r["d_in"][i]), int(r["d_out"][j]), int(r["a2"][i, j]), int(r["sharp"][i, j]), int(r["lam"][i, j])) \
        == (di, dj, a2, sharp, lam)
    assert abs(float(r["C"][i, j]) - c) < 2e-6
    p = bfc_paper(ei, n, np.array([[i, j]]))
    tri, s1, s2, gamma, val = paper
    assert (int(p["tri"][0]), int(p["sq_i"][0]), int(p["sq_j"][0]), int(p["gamma"][0])) == (tri, s1, s2, gamma)
    assert abs(float(p["bfc"][0]) - val) < 1e-12


def test_k33_compiled_dataflow_sign():
    # the compiled kernel's two-rounding dataflow gives -1.9868216e-08 on K3,3 (simulator: > 0)
    assert float(closing_value(3, 3, 0, 1, 6, 3, "compiled")) == pytest.approx(-1.9868216e-08, rel=1e-6)
    assert float(closing_value(3, 3, 0, 1, 6, 3, "sim32")) > 0


@pytest.mark.parametrize("seed", range(6))
def test_closed_form_identity_of_cuda_flavour(seed):
    # lambda == d_max and sharp == d_i + d_j - t1 - t2 for symmetric 0/1 A without self-loops
    n = 12 + 3 * seed
    ei = gnp(n, 0.25, seed)
    A = dense_of(ei, n)
    r = bfc_cuda_dense(A)
    A2 = r["a2"]
    deg = A.sum(1)
    for i, j in zip(*np.nonzero(A)):
        tri_nodes = np.flatnonzero((A[i] > 0) & (A[j] > 0))
        t1 = sum(A2[i, k] == 1 for k in tri_nodes)
        t2 = sum(A2[k, j] == 1 for k in tri_nodes)
        assert r["sharp"][i, j] == deg[i] + deg[j] - t1 - t2
        assert r["lam"][i, j] == max(deg[i], deg[j])
    assert np.array_equal(r["C"], r["C"].T)


def test_choice_index_matches_numpy_choice():
    rng = np.random.default_rng(0)
    for trial in range(300):
        n = int(rng.integers(1, 40))
        a = rng.normal(size=n).astype(np.float32).astype(np.float64)
        tau = [float("inf"), 1, 12, 50][trial % 4]
        p = softmax(a, tau)
        seed = int(rng.integers(0, 2**31))
        np.random.seed(seed)
        want = np.random.choice(range(n), p=p)
        u = np.random.RandomState(seed).random_sample()
        assert choice_index(p, u) == want


def test_softmax_overflow_raises_like_numpy():
    p = softmax(np.array([2.0, 1.0]), tau=436)
    assert np.isnan(p).any()
    with pytest.raises(ValueError, match="NaN"):
        choice_index(p, 0.5)


def test_c_oracle_agrees_with_python_oracle():
    from dcr import graph
    from dcr.synth import named_graph
    from oracle.c_port import bfc_paper_c
    from oracle.paper_flavour import bfc_paper
    ei, n = named_graph("wisconsin")
    rowptr, col = graph.undirected_csr(ei, n)
    ref = bfc_paper(ei, n)
    c = bfc_paper_c(rowptr, col, ref["edges"][:, 0], ref["edges"][:, 1], threads=2)
    for k in ("tri", "sq_i", "sq_j", "gamma", "bfc"):
        assert np.array_equal(c[k], ref[k]), k


def test_reference_kernel_ptx_recipe_when_reference_is_present():
    """oracle/build_ref.py: the reference's two numba kernels compile to PTX without a GPU (build container only —
    /root/reference does not exist on the GPU box, where the prebuilt oracle/_ref travels instead)."""
    import json
    import os
    import pytest
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference checkout not present")
    from oracle import build_ref
    assert build_ref.build()
    with open(os.path.join(build_ref.OUT, "manifest.json")) as f:
        m = json.load(f)
    k = m["kernels"]
    # numba's flattened Array ABI: 5 + 2*ndim params per array (oracle/ref_gpu.py builds exactly these)
    assert k["_balanced_forman_curvature"]["n_params"] == 9 + 9 + 7 + 7 + 1 + 9
    assert k["_balanced_forman_post_delta"]["n_params"] == 9 + 9 + 1 + 1 + 1 + 9 + 1 + 1 + 7 + 7 + 1 + 1
    for v in k.values():
        ptx = open(os.path.join(build_ref.OUT, v["file"])).read()
        assert v["entry"] in ptx and ".target sm_90" in ptx

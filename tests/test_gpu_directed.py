"""GPU parity, directed mode (is_undirected=False): the definitional kernels for an asymmetric 0/1
adjacency (csrc/dcr_directed.cuh) and the directed SDRF loop against the oracle (compiled rounding model), the goldens
of the unmodified reference and the reference's own compiled numba kernels run on this GPU."""
import numpy as np
import pytest

from helpers import golden

pytestmark = pytest.mark.gpu


def random_digraph(n, m, seed, shuffle=True):
    rng = np.random.default_rng(seed)
    ei = rng.integers(0, n, size=(2, m))
    ei = ei[:, ei[0] != ei[1]]
    ei = np.unique(ei, axis=1)
    if shuffle:
        ei = ei[:, rng.permutation(ei.shape[1])]
    return ei


def dense_directed(ei, n):
    A = np.zeros((n, n), dtype=np.float32)
    A[ei[0], ei[1]] = 1
    return A


def _oracle_log_tuples(log):
    return [(r["x"], r["y"], r["n_candidates"], r["k"], r["l"], r["choice"],
             -1 if r["removed"] is None else r["removed"][0], -1 if r["removed"] is None else r["removed"][1])
            for r in log]


@pytest.mark.parametrize("n,m,seed", [(12, 30, 0), (40, 300, 1), (64, 900, 2), (300, 2500, 3), (1500, 9000, 4)])
def test_directed_curvature_bit_identical_to_oracle(n, m, seed):
    """balanced_forman_curvature(A) with asymmetric A: fp32 image of all N^2 entries, plus the integer fields per entry."""
    import torch
    from curvature.bfc_cuda import balanced_forman_curvature
    from dcr import bfc
    from oracle.cuda_flavour import bfc_cuda_dense
    ei = random_digraph(n, m, seed)
    An = dense_directed(ei, n)
    ref = bfc_cuda_dense(An)
    C = torch.full((n, n), 3.0, device="cuda")
    out = balanced_forman_curvature(torch.from_numpy(An).cuda(), C=C)
    assert out is C
    assert np.array_equal(C.cpu().numpy().view(np.uint32), ref["C"].view(np.uint32))
    d = bfc.DirectedCSR.from_edge_index(ei, n)
    f = bfc.cuda_flavour_directed(d)
    rows = np.repeat(np.arange(n), np.diff(d.out.rowptr.cpu().numpy()))
    cols = d.out.colidx.cpu().numpy()
    A2 = An @ An
    assert np.array_equal(f["tri"].cpu().numpy(), A2[rows, cols].astype(np.int32))
    if "sharp" in ref:
        assert np.array_equal(f["sharp"].cpu().numpy(), ref["sharp"][rows, cols].astype(np.int32))
        assert np.array_equal(f["lam"].cpu().numpy(), ref["lam"][rows, cols].astype(np.int32))


def test_directed_curvature_zero_degree_entries_and_reciprocal_pairs():
    """d_in[i] = 0 or d_out[j] = 0 -> C = 0 (bfc_cuda.py:27-29); reciprocal pairs and 2-cycles exercise A2[i,i]."""
    import torch
    from curvature.bfc_cuda import balanced_forman_curvature
    from oracle.cuda_flavour import bfc_cuda_dense
    pairs = [(0, 1), (1, 2), (2, 0), (0, 2), (3, 0), (3, 1), (1, 3), (4, 5), (5, 4), (6, 4)]
    n = 7
    An = dense_directed(np.array(pairs).T, n)
    got = balanced_forman_curvature(torch.from_numpy(An).cuda()).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), bfc_cuda_dense(An)["C"].view(np.uint32))
    assert An[6, 4] == 1 and got[6, 4] == 0            # node 6 has no predecessor: d_in[6] = 0
    assert got[4, 5] != 0


@pytest.mark.parametrize("n,m,seed", [(14, 40, 5), (30, 200, 6), (60, 500, 7)])
def test_directed_post_delta_bit_identical_to_oracle(n, m, seed):
    import torch
    from curvature.bfc_cuda import balanced_forman_post_delta
    from oracle.cuda_flavour import post_delta_dense
    ei = random_digraph(n, m, seed)
    An = dense_directed(ei, n)
    A = torch.from_numpy(An).cuda()
    rng = np.random.default_rng(seed)
    cols = rng.choice(ei.shape[1], size=6, replace=False)
    pairs = [(int(ei[0][c]), int(ei[1][c])) for c in cols] + [(0, 0), (1, 1)]
    pairs += [(int(a), int(b)) for a, b in rng.integers(0, n, size=(3, 2))]       # arbitrary (x, y), mostly non-edges
    for x, y in pairs:
        xn = rng.permutation(np.flatnonzero(An[x])).tolist() + [x]                 # successors of x (:48)
        yn = rng.permutation(np.flatnonzero(An[:, y])).tolist() + [y]              # predecessors of y (:49)
        want = post_delta_dense(An, x, y, xn, yn, "compiled")
        got = balanced_forman_post_delta(A, x, y, xn, yn).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (x, y)
    # arbitrary index lists (not neighbour lists, with repeats)
    x, y = int(ei[0][0]), int(ei[1][0])
    xn = rng.integers(0, n, size=9).tolist()
    yn = rng.integers(0, n, size=7).tolist()
    want = post_delta_dense(An, x, y, xn, yn, "compiled")
    got = balanced_forman_post_delta(A, x, y, xn, yn).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def _run_both(ei, n, loops, bound, tau, seed, remove_edges=True):
    from dcr import sdrf
    from oracle.sdrf import sdrf_oracle
    uni = np.random.RandomState(seed).random_sample(loops)
    got, glog = sdrf.sdrf(ei, n, loops, remove_edges, bound, tau, uniforms=uni, return_log=True, is_undirected=False)
    want, wlog = sdrf_oracle(ei, n, loops, remove_edges, bound, tau, uni, rounding="compiled", is_undirected=False)
    assert [tuple(int(v) for v in r) for r in glog] == _oracle_log_tuples(wlog)
    assert np.array_equal(got, want)
    return glog


@pytest.mark.parametrize("tau", [float("inf"), 5, 40])
@pytest.mark.parametrize("seed", range(4))
def test_directed_sdrf_matches_oracle(seed, tau):
    n = 12 + 6 * seed
    ei = random_digraph(n, 3 * n + 10 * seed, 100 + seed)
    _run_both(ei, n, 14, [0.3, 0.5, 0.1, 0.8][seed], tau, 60 + seed)


def test_directed_sdrf_without_removal_and_larger_graph():
    ei = random_digraph(120, 700, 9)
    _run_both(ei, 120, 40, 0.4, 10, 3)
    _run_both(ei, 120, 25, 0.4, float("inf"), 3, remove_edges=False)


def test_directed_sdrf_incremental_refresh_equals_full_recompute():
    """After the loop, the maintained curvatures equal a fresh evaluation of the final graph (dirty-set completeness)."""
    import torch
    from dcr import bfc, sdrf
    n = 80
    ei = random_digraph(n, 500, 11)
    loops = 60
    uni = np.random.RandomState(1).random_sample(loops)
    keep = []
    out = sdrf.sdrf(ei, n, loops, True, 0.3, 8, uniforms=uni, is_undirected=False, state_out=keep)
    st = keep[0]
    rowptr, order, col, c32, _ = st.export(with_curvature=True)
    st.close()
    d = bfc.DirectedCSR.from_edge_index(out, n)
    fresh = bfc.cuda_flavour_directed(d, want_fields=False)["c32"].cpu().numpy()
    assert np.array_equal(col, d.out.colidx.cpu().numpy())
    assert np.array_equal(c32.view(np.uint32), fresh.view(np.uint32))


def test_directed_sdrf_golden_sequences_of_the_unmodified_reference():
    """tests/golden/sdrf_directed_seq.npz was produced under the numba simulator (all-fp32 closing formula); the cases
    where that rounding model and the compiled one agree must reproduce its add/remove sequence and output."""
    from dcr import sdrf
    from oracle.sdrf import sdrf_oracle
    z = golden("sdrf_directed_seq.npz")
    agreeing = 0
    for name in (str(s) for s in z["names"]):
        ei, n = z[f"{name}/edge_index"], int(z[f"{name}/n"])
        args = (int(z[f"{name}/loops"]), True, float(z[f"{name}/bound"]), float(z[f"{name}/tau"]))
        uni = z[f"{name}/uniforms"]
        _, wlog = sdrf_oracle(ei, n, *args, uni, rounding="compiled", is_undirected=False)
        seq_c = []
        for r in wlog:
            if r["k"] >= 0:
                seq_c.append((1, r["k"], r["l"]))
            if r["removed"] is not None:
                seq_c.append((-1,) + tuple(r["removed"]))
        if not np.array_equal(np.array(seq_c, dtype=np.int64).reshape(-1, 3), z[f"{name}/log"]):
            continue          # a near-tie the two rounding models break differently
        agreeing += 1
        got, log = sdrf.sdrf(ei, n, *args, uniforms=uni, return_log=True, is_undirected=False)
        seq = []
        for r in log:
            if r[3] >= 0:
                seq.append((1, int(r[3]), int(r[4])))
            if r[6] >= 0:
                seq.append((-1, int(r[6]), int(r[7])))
        assert np.array_equal(np.array(seq, dtype=np.int64).reshape(-1, 3), z[f"{name}/log"]), name
        assert np.array_equal(got, z[f"{name}/out"]), name
    assert agreeing >= 3


def test_directed_against_the_reference_kernels_on_this_gpu():
    """The reference's own compiled numba kernels (oracle/_ref PTX) on asymmetric A, and its loop driven with them."""
    import torch
    from oracle import ref_gpu
    if not ref_gpu.available():
        pytest.skip("oracle/_ref PTX not built")
    from curvature.bfc_cuda import balanced_forman_curvature
    from dcr import sdrf
    for n, m, seed in [(20, 90, 21), (50, 400, 22)]:
        ei = random_digraph(n, m, seed)
        A = torch.from_numpy(dense_directed(ei, n)).cuda()
        ref = ref_gpu.balanced_forman_curvature(A)
        got = balanced_forman_curvature(A)
        assert torch.equal(ref.view(torch.int32), got.view(torch.int32))
        loops = 12
        uni = np.random.RandomState(seed).random_sample(loops)
        want, wlog = ref_gpu.sdrf_reference_gpu(ei, n, loops, True, 0.3, 7, uni, is_undirected=False)
        out, log = sdrf.sdrf(ei, n, loops, True, 0.3, 7, uniforms=uni, return_log=True, is_undirected=False)
        assert [tuple(int(v) for v in r) for r in log] == _oracle_log_tuples(wlog)
        assert np.array_equal(out, want)


def test_directed_dropin_signature_and_rejections():
    import torch
    from rewiring.sdrf_cuda_bfc import sdrf_cuda_bfc
    from torch_geometric.data import Data
    from oracle.sdrf import sdrf_oracle
    n = 25
    ei = random_digraph(n, 110, 31)
    data = Data(edge_index=torch.from_numpy(ei).long())
    data.num_nodes = n
    uni = np.random.RandomState(4).random_sample(10)
    out = sdrf_cuda_bfc(data, 10, True, 0.4, 6, False, uniforms=uni)
    want, _ = sdrf_oracle(ei, n, 10, True, 0.4, 6, uni, rounding="compiled", is_undirected=False)
    assert np.array_equal(out.edge_index.numpy(), want)
    dup = np.concatenate([ei, ei[:, :1]], axis=1)
    data = Data(edge_index=torch.from_numpy(dup).long())
    data.num_nodes = n
    with pytest.raises(NotImplementedError):
        sdrf_cuda_bfc(data, 3, True, 0.4, 6, False, uniforms=uni)


def test_directed_sdrf_inputs_with_self_loops():
    """Self-loops stay in the reference's DiGraph (a node is its own successor and predecessor) but not in A: oracle,
    goldens of the unmodified reference, and a negative bound that lets the (0,0) fallback delete the loop 0 -> 0."""
    from dcr import sdrf
    from oracle.sdrf import sdrf_oracle
    rng = np.random.default_rng(3)
    for s, tau in enumerate([float("inf"), 7]):
        n = 14 + 8 * s
        ei = random_digraph(n, 4 * n, 200 + s)
        who = rng.choice(n, size=4 + s, replace=False)
        ei = np.concatenate([ei, np.stack([who, who])], axis=1)
        ei = ei[:, rng.permutation(ei.shape[1])]
        _run_both(ei, n, 12, 0.3, tau, 90 + s)
    n = 8
    ring = np.array([[i, (i + 1) % n] for i in range(n)]).T
    ei = np.concatenate([ring, np.stack([np.arange(n), np.arange(n)])], axis=1)      # a directed ring: every C is <= 0
    _run_both(ei, n, 4, -5.0, float("inf"), 1)
    z = golden("sdrf_directed_selfloop_seq.npz")
    agreeing = 0
    for name in (str(s) for s in z["names"]):
        ei, n = z[f"{name}/edge_index"], int(z[f"{name}/n"])
        args = (int(z[f"{name}/loops"]), True, float(z[f"{name}/bound"]), float(z[f"{name}/tau"]))
        uni = z[f"{name}/uniforms"]
        want, _ = sdrf_oracle(ei, n, *args, uni, rounding="compiled", is_undirected=False)
        if not np.array_equal(want, z[f"{name}/out"]):
            continue
        agreeing += 1
        got = sdrf.sdrf(ei, n, *args, uniforms=uni, is_undirected=False)
        assert np.array_equal(got, z[f"{name}/out"]), name
    assert agreeing >= 2

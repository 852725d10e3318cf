"""Small inputs through every kernel added in round 2 — run under compute-sanitizer (memcheck / racecheck):

    compute-sanitizer --tool memcheck python profiles/sanitizer_driver.py
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "discrete-curvature-rewiring_b200"))
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, REPO)
from dcr import bfc, graph, sdrf  # noqa: E402
from dcr import dist as ddist  # noqa: E402
from dcr.synth import chung_lu_graph, named_graph  # noqa: E402
from helpers import gnp, sym_edge_index  # noqa: E402

rng = np.random.default_rng(0)
# directed kernels + loop
n = 40
dei = rng.integers(0, n, size=(2, 260))
dei = np.unique(dei[:, dei[0] != dei[1]], axis=1)
dei = dei[:, rng.permutation(dei.shape[1])]
d = bfc.DirectedCSR.from_edge_index(dei, n)
bfc.cuda_flavour_directed(d)
A = torch.zeros(n, n, device="cuda")
A[torch.from_numpy(dei[0]).cuda(), torch.from_numpy(dei[1]).cuda()] = 1
from curvature.bfc_cuda import balanced_forman_curvature, balanced_forman_post_delta  # noqa: E402
balanced_forman_curvature(A)
x, y = int(dei[0][0]), int(dei[1][0])
xn = np.flatnonzero(A[x].cpu().numpy()).tolist() + [x]
yn = np.flatnonzero(A[:, y].cpu().numpy()).tolist() + [y]
balanced_forman_post_delta(A, x, y, xn, yn)
uni = np.random.RandomState(1).random_sample(30)
sdrf.sdrf(dei, n, 30, True, 0.3, 5, uniforms=uni, is_undirected=False)
# classical loops
ei = gnp(50, 0.15, 3)
for ct in ("1d", "augmented", "haantjes"):
    sdrf.sdrf(ei, 50, 30, True, 0.5 if ct != "1d" else -4.0, 2, uniforms=uni, curv_type=ct)
# undirected BFC loop (fused shifts, one-pass argmin) incl. row relocation
sdrf.sdrf(sym_edge_index([(0, i) for i in range(1, 30)], 30), 30, 30, True, 0.5, float("inf"), uniforms=uni)
# edge-centric cuda flavour incl. a hub-hub edge (>= 512 common-row entries) and the sharded entry point at world 1
hub = sym_edge_index([(0, i) for i in range(1, 700)] + [(1, i) for i in range(2, 650)], 700)
for e2, n2 in ((hub, 700), (ei, 50)):
    rowptr, col = graph.undirected_csr(e2, n2)
    csr = bfc.DeviceCSR.from_host(rowptr, col)
    bfc.cuda_flavour_edges(csr)
    sc = ddist.ShardedCudaBFC(csr)
    sc.run()
    torch.cuda.synchronize()
    sc.close()
# symmetric tensor-core product + mirror pass, large post_delta matrix (several CTAs)
e3 = chung_lu_graph(300, 6000, 0.7, 0.3, 2)
rowptr, col = graph.undirected_csr(e3, 300)
csr = bfc.DeviceCSR.from_host(rowptr, col)
bfc.support_tc(csr)
bfc.cuda_flavour_tc(csr)
tri = bfc.support(csr)
deg = np.diff(rowptr)
x = int(np.argmax(deg))
y = int(col[rowptr[x]])
i_nb = torch.from_numpy(np.append(col[rowptr[x]:rowptr[x + 1]], x).astype(np.int32)).cuda()
j_nb = torch.from_numpy(np.append(col[rowptr[y]:rowptr[y + 1]], y).astype(np.int32)).cuda()
D = torch.zeros(i_nb.numel(), j_nb.numel(), device="cuda")
bfc.post_delta(csr, tri, x, y, i_nb, j_nb, D)
# paper flavour with the round-2 planning (tiers, stream-bounded runs) on a graph with hubs, both modes
from dcr import lib as L  # noqa: E402
e4, n4 = named_graph("cora")
rowptr, col = graph.undirected_csr(e4, n4)
csr = bfc.DeviceCSR.from_host(rowptr, col)
for mode in (0, 1):
    L.load().dcr_bfc_paper_set_mode(mode)
    bfc.paper_flavour(csr)
L.load().dcr_bfc_paper_set_mode(0)
torch.cuda.synchronize()
print("sanitizer driver done")

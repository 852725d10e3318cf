"""Calibrate the per-edge work estimate behind the contiguous work-balanced ranges (dcr_bfc_paper_edge_cost).

Measures, on ONE GPU, the edge-kernel time of many contiguous ranges of the arxiv-shaped edge list (the W = 4, 6, 8, 12
cuts of the current estimate) and fits  t = wL * stream(L0 edges) + wG * stream(group-class warp edges) + wC *
stream(cooperative edges) + wH * heads + wE * edges + t0  by non-negative least squares.  The ratios go into
edge_cost_kernel (csrc/dcr_bfc_paper.cu)."""
import os
import sys

import numpy as np
import torch
from scipy.optimize import nnls

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "discrete-curvature-rewiring_b200"))
sys.path.insert(0, REPO)
from dcr import bfc, graph  # noqa: E402
from dcr import dist as ddist  # noqa: E402
from dcr import lib as L  # noqa: E402
from dcr.synth import named_graph  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "arxiv"
ei, n = named_graph(name)
rowptr, col = graph.undirected_csr(ei, n)
csr = bfc.DeviceCSR.from_host(rowptr, col)
esrc, edst, _ = csr.undirected_edges()
E = int(esrc.numel())
deg = np.diff(rowptr).astype(np.int64)
es, ed = esrc.cpu().numpy(), edst.cpu().numpy()
S = np.add.reduceat(deg[col], rowptr[:-1].astype(np.int64)) * (deg > 0)
ca, cb = S[ed] - deg[es], S[es] - deg[ed]
sw = cb < ca
stream = np.where(sw, cb, ca)
da = np.where(sw, deg[ed], deg[es])
heads = np.where(sw, deg[es], deg[ed])
triv = np.minimum(deg[es], deg[ed]) <= 1
coop = (stream > 16384) & ~triv
l0 = (da <= 128) & ~coop & ~triv
grp = (da > 128) & ~coop & ~triv
F = np.stack([stream * l0, stream * grp, stream * coop, heads * ~triv, np.ones(E), stream * grp * da, stream * coop * da],
             axis=1).astype(np.float64)
P = np.concatenate([np.zeros((1, F.shape[1])), np.cumsum(F, axis=0)])
lib = L.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
pre = torch.cumsum(ddist.edge_cost(csr, esrc, edst), 0)
comm = ddist.PeerComm(E, 0, 1)
nbytes = int(lib.dcr_bfc_paper_scratch_bytes(n, csr.max_degree, E))
scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
rows, times = [], []
cuts = [ddist.balanced_bounds(pre, w) for w in (4, 8)]
cuts += [[(E * k) // w for k in range(w + 1)] for w in (6, 10, 16)]          # equal COUNTS: very different mixes
for b in cuts:
    world = len(b) - 1
    for r in range(world):
        lo, cnt = b[r], b[r + 1] - b[r]
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for a_, c_ in ev:
            a_.record(); c_.record()
        for k in range(-2, 5):
            flush.zero_()
            e0, e1 = (ev[k][0].cuda_event, ev[k][1].cuda_event) if k >= 0 else (0, 0)
            L.check(lib.dcr_bfc_paper_sharded(csr.rowptr.data_ptr(), csr.colidx.data_ptr(), n, csr.max_degree, esrc.data_ptr(),
                                              edst.data_ptr(), lo, cnt, comm.handle, scratch.data_ptr(), nbytes, e0, e1,
                                              L.current_stream()), "sharded")
        torch.cuda.synchronize()
        times.append(float(np.median([a_.elapsed_time(c_) for a_, c_ in ev])))
        rows.append(np.concatenate([P[lo + cnt] - P[lo], [1.0]]))
A, t = np.array(rows), np.array(times)
scale = A.max(axis=0)
w, res = nnls(A / scale, t)
w = w / scale
pred = A @ w
print(f"{name}: {len(t)} ranges; fit (ms per unit): wL {w[0]:.3e} wG {w[1]:.3e} wC {w[2]:.3e} wH {w[3]:.3e} wE {w[4]:.3e} "
      f"wG*da {w[5]:.3e} wC*da {w[6]:.3e} t0 {w[7]:.3f} ms")
w = np.maximum(w, 1e-30)
print(f"relative to wL: wG {w[1] / w[0]:.2f} wC {w[2] / w[0]:.2f} wH {w[3] / w[0]:.1f} wE {w[4] / w[0]:.1f}; "
      f"rms residual {np.sqrt(np.mean((pred - t) ** 2)) * 1e3:.1f} us, max {np.abs(pred - t).max() * 1e3:.1f} us")
print("W=8 measured", np.round(t[4:12], 3), "predicted", np.round(pred[4:12], 3))
# the same fit without the constant and with wL pinned by the light kernel's own rate (sum of work, what a cut balances)
for cols, label in (([0, 1, 2, 5, 6], "L G C G*da C*da"), ([0, 1, 2], "L G C")):
    X = A[:, cols]
    sc2 = X.max(axis=0)
    w2, _ = nnls(X / sc2, t - 0.15)
    w2 = w2 / sc2
    print(label, "no constant (t - 0.15 ms):", " ".join(f"{v:.3e}" for v in w2), "rms us", round(float(np.sqrt(np.mean((X @ w2 + 0.15 - t) ** 2))) * 1e3, 1))
np.save(os.path.join(REPO, "gpurun_out", "cost_fit_A.npy"), np.column_stack([A, t]))

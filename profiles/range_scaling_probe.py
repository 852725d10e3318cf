"""Time EVERY rank's contiguous work-balanced range of the arxiv-shaped pass for several world sizes on ONE GPU (no
exchange): per-rank step and edge-kernel times, i.e. what the slowest rank of a W-GPU run would take — shows the balance
of the work estimate and how the per-rank kernel time scales when the range shrinks.

    python profiles/range_scaling_probe.py          # PROBE_WORLDS=1,2,4,8 by default
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "discrete-curvature-rewiring_b200"))
sys.path.insert(0, REPO)
from dcr import bfc, graph  # noqa: E402
from dcr import dist as ddist  # noqa: E402
from dcr import lib as L  # noqa: E402
from dcr.synth import named_graph  # noqa: E402

ei, n = named_graph("arxiv")
rowptr, col = graph.undirected_csr(ei, n)
csr = bfc.DeviceCSR.from_host(rowptr, col)
esrc, edst, _ = csr.undirected_edges()
E = int(esrc.numel())
lib = L.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
cost = ddist.edge_cost(csr, esrc, edst)
pre = torch.cumsum(cost, 0)
comm = ddist.PeerComm(E, 0, 1)
nbytes = int(lib.dcr_bfc_paper_scratch_bytes(n, csr.max_degree, E))
scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
worlds = [int(w) for w in os.environ.get("PROBE_WORLDS", "1,2,4,8").split(",")]
for world in worlds:
    b = ddist.balanced_bounds(pre, world)
    steps, edges = [], []
    for r in range(world):
        lo, cnt = b[r], b[r + 1] - b[r]
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(6)]
        st = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(6)]
        for a, c in ev:
            a.record(); c.record()

        def one(k=None):
            e0 = ev[k][0].cuda_event if k is not None else 0
            e1 = ev[k][1].cuda_event if k is not None else 0
            L.check(lib.dcr_bfc_paper_sharded(csr.rowptr.data_ptr(), csr.colidx.data_ptr(), n, csr.max_degree, esrc.data_ptr(),
                                              edst.data_ptr(), lo, cnt, comm.handle, scratch.data_ptr(), nbytes, e0, e1,
                                              L.current_stream()), "sharded")
        for _ in range(2):
            one()
        for k in range(6):
            flush.zero_()
            st[k][0].record()
            one(k)
            st[k][1].record()
        torch.cuda.synchronize()
        steps.append(float(np.median([a.elapsed_time(c) for a, c in st])))
        edges.append(float(np.median([a.elapsed_time(c) for a, c in ev])))
    print(f"world {world}: step max {max(steps):.3f} mean {np.mean(steps):.3f} ms | edge kernels max {max(edges):.3f} "
          f"mean {np.mean(edges):.3f} | per rank {[round(x, 3) for x in edges]} | edges {[b[r+1]-b[r] for r in range(world)]}", flush=True)

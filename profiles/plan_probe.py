"""Times the graph-static part of a pass: dcr_bfc_paper_edge_cost = node_s_kernel + edge_cost_kernel (events, L2 flushed)."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "discrete-curvature-rewiring_b200"))
from dcr import bfc, graph  # noqa: E402
from dcr import dist as ddist  # noqa: E402
from dcr.synth import named_graph  # noqa: E402

ei, n = named_graph("arxiv")
rowptr, col = graph.undirected_csr(ei, n)
csr = bfc.DeviceCSR.from_host(rowptr, col)
esrc, edst, _ = csr.undirected_edges()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for k in range(12):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ddist.edge_cost(csr, esrc, edst)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(os.environ.get("DCR_LIB_PATH", "default"), "edge_cost (node_s + cost kernels + 2 allocations): median", round(float(np.median(ts[2:])) * 1e3, 1), "us")

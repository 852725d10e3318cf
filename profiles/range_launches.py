"""One W = 8 range (PROBE_RANK, default 7) of the arxiv-shaped pass, a few passes — run under
`ncu --metrics gpu__time_duration.sum` to list what a 1/8 step launches and how long each kernel takes."""
import os
import sys

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
world, r = int(os.environ.get("PROBE_WORLD", "8")), int(os.environ.get("PROBE_RANK", "7"))
b = ddist.balanced_bounds(torch.cumsum(ddist.edge_cost(csr, esrc, edst), 0), world)
comm = ddist.PeerComm(E, 0, 1)
nbytes = int(lib.dcr_bfc_paper_scratch_bytes(n, csr.max_degree, E))
scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
for _ in range(int(os.environ.get("PROBE_PASSES", "4"))):
    L.check(lib.dcr_bfc_paper_sharded(csr.rowptr.data_ptr(), csr.colidx.data_ptr(), n, csr.max_degree, esrc.data_ptr(),
                                      edst.data_ptr(), b[r], b[r + 1] - b[r], comm.handle, scratch.data_ptr(), nbytes, 0, 0,
                                      L.current_stream()), "sharded")
torch.cuda.synchronize()
print("ok")

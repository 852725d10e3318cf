"""Times the cuda-flavour CSR routes on the arxiv- and squirrel-shaped graphs (one GPU): the per-entry kernels of
dcr_bfc_cuda.cu (supports + closing, every edge twice) against the edge-centric kernels of dcr_bfc_cuda_edges.cu."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "discrete-curvature-rewiring_b200"))
from dcr import bfc, graph  # noqa: E402
from dcr.synth import named_graph  # noqa: E402


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for r in range(reps):
        flush.zero_()
        ev[r].record()
        fn()
        ev[r + 1].record()
        torch.cuda.synchronize()
        tot += ev[r].elapsed_time(ev[r + 1])
    return tot / reps


for name in ("arxiv", "squirrel"):
    ei, n = named_graph(name)
    rowptr, col = graph.undirected_csr(ei, n)
    csr = bfc.DeviceCSR.from_host(rowptr, col)
    e = csr.nnz // 2
    tri = bfc.support(csr)
    out = bfc.cuda_flavour_edges(csr)
    t_sup = timed(lambda: bfc.support(csr, out=tri))
    t_all = timed(lambda: bfc.cuda_flavour(csr, want_fields=False, tri=bfc.support(csr, out=tri)))   # tri given: per-entry kernels
    t_e1 = timed(lambda: bfc.cuda_flavour_edges(csr, phases=1, out=out))
    t_e = timed(lambda: bfc.cuda_flavour_edges(csr, out=out))
    print(f"{name}: E={e} per-entry support {t_sup:.3f} ms, per-entry full {t_all:.3f} ms | edge-centric support "
          f"{t_e1:.3f} ms, full {t_e:.3f} ms -> {e / t_e / 1e3:.1f} M edges/s")

"""Per-CTA timeline of the paper-flavour group kernel (which work item held a CTA longest, when each phase ended).

Needs a trace build of the library (the hooks compile to nothing otherwise):

    cd discrete-curvature-rewiring_b200/csrc
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
         -DDCR_PAPER_TRACE -o /tmp/libdcr_trace.so *.cu
    DCR_LIB_PATH=/tmp/libdcr_trace.so python profiles/paper_cta_timeline.py        # on a B200

Prints, for rank 0's share of the arxiv-shaped pass at world sizes 1 and 8: the kernel span, when the cooperative
phase and the CTAs ended (median / p90 / max), and the eight longest single work items with the edge behind them
(phase 3 = part of a split edge, 1 = cooperative edge, 2 = run of light edges).  This is how the round-1 scaling floor
(single hub-hub edges holding one CTA for 3-4 ms) was found.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "discrete-curvature-rewiring_b200"))
sys.path.insert(0, REPO)
from dcr import bfc, graph, lib as L  # noqa: E402
from dcr.synth import named_graph  # noqa: E402

ei, n = named_graph("arxiv")
rowptr, col = graph.undirected_csr(ei, n)
csr = bfc.DeviceCSR.from_host(rowptr, col)
esrc, edst, _ = csr.undirected_edges()
E = int(esrc.numel())
deg = np.diff(rowptr)
lib = L.load()
if not hasattr(lib, "dcr_paper_trace_read"):
    raise SystemExit("this libdcr.so was built without -DDCR_PAPER_TRACE (see the docstring)")
lib.dcr_paper_trace_read.restype = C.c_int
for world in (1, 8):
    count = bfc.shard_count(E, 0, world)
    ws = bfc.PaperWorkspace(csr, count, chunk=max(1, (E + world - 1) // world))
    for _ in range(3):
        bfc.paper_flavour(csr, rank=0, world=world, ws=ws)
    torch.cuda.synchronize()
    nct = torch.cuda.get_device_properties(0).multi_processor_count * 4
    buf = np.zeros(nct * 6, dtype=np.uint64)
    lib.dcr_paper_trace_read(C.c_void_p(buf.ctypes.data), C.c_int(nct))
    t = buf.reshape(nct, 6).astype(np.int64)
    t0 = t[:, 0].min()
    us = lambda x: (x - t0) / 1e3
    print(f"world {world}: kernel span {us(t[:, 2].max()):.0f} us; cooperative phase ends: median {us(np.median(t[:, 1])):.0f} "
          f"max {us(t[:, 1].max()):.0f} us; CTA ends: median {us(np.median(t[:, 2])):.0f} p90 {us(np.quantile(t[:, 2], .9)):.0f} "
          f"max {us(t[:, 2].max()):.0f} us")
    es, ed = esrc.cpu().numpy(), edst.cpu().numpy()
    for c in np.argsort(-t[:, 3])[:8]:
        phase = int(t[c, 4] >> 30)
        extra = ""
        if phase in (1, 3):
            e = int(t[c, 4] & ((1 << 30) - 1)) * world
            extra = f"edge ({es[e]},{ed[e]}) degrees ({deg[es[e]]},{deg[ed[e]]})"
        print(f"   cta {c}: longest item {t[c, 3] / 1e3:.0f} us, phase {phase}, {t[c, 5]} items, ended at {us(t[c, 2]):.0f} us {extra}")

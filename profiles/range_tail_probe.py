"""Where does a 1/8 range of the arxiv-shaped pass spend its time?  For every rank's contiguous work-balanced range at
W = 8 (on ONE GPU): edge-kernel time of the library given by DCR_LIB_PATH (builds with -DDCR_ONLY_GROUP / -DDCR_ONLY_LIGHT
time one kernel alone), and with a -DDCR_PAPER_TRACE build the per-CTA timeline of the group kernel.

    DCR_LIB_PATH=build/libdcr_trace.so python profiles/range_tail_probe.py
"""
import ctypes as C
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
deg = np.diff(rowptr)
lib = L.load()
trace = hasattr(lib, "dcr_paper_trace_read")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
cost = ddist.edge_cost(csr, esrc, edst)
pre = torch.cumsum(cost, 0)
comm = ddist.PeerComm(E, 0, 1)
nbytes = int(lib.dcr_bfc_paper_scratch_bytes(n, csr.max_degree, E))
scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
world = int(os.environ.get("PROBE_WORLD", "8"))
b = ddist.balanced_bounds(pre, world)
es, ed = esrc.cpu().numpy(), edst.cpu().numpy()
S = np.add.reduceat(deg[col], rowptr[:-1].astype(np.int64)) * (deg > 0)
stream = np.minimum(S[ed] - deg[es], S[es] - deg[ed])
da = np.where(S[es] - deg[ed] < S[ed] - deg[es], deg[ed], deg[es])
print(os.environ.get("DCR_LIB_PATH", "default lib"), "trace" if trace else "")
for r in range(world):
    lo, cnt = b[r], b[r + 1] - b[r]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(6)]
    for a_, c_ in ev:           # torch only lets elapsed_time() run on events it has recorded itself once
        a_.record(); c_.record()

    def one(k=None):
        e0 = ev[k][0].cuda_event if k is not None else 0
        e1 = ev[k][1].cuda_event if k is not None else 0
        L.check(lib.dcr_bfc_paper_sharded(csr.rowptr.data_ptr(), csr.colidx.data_ptr(), n, csr.max_degree, esrc.data_ptr(),
                                          edst.data_ptr(), lo, cnt, comm.handle, scratch.data_ptr(), nbytes, e0, e1,
                                          L.current_stream()), "sharded")
    for _ in range(2):
        one()
    if trace and hasattr(lib, "dcr_paper_trace_counters"):
        torch.cuda.synchronize()
        cz = np.zeros(8, dtype=np.uint64)
        lib.dcr_paper_trace_counters(C.c_void_p(cz.ctypes.data), C.c_int(1))
    for k in range(6):
        flush.zero_()
        one(k)
    torch.cuda.synchronize()
    if trace and hasattr(lib, "dcr_paper_trace_counters"):
        lib.dcr_paper_trace_counters(C.c_void_p(cz.ctypes.data), C.c_int(1))
        print(f"    per pass: deferred edges {int(cz[0]) // 6}, CTA-path edges {int(cz[3]) // 6} of which redone with the global hash {int(cz[2]) // 6}; phase-3 items {int(cz[5]) // 6}, mean {int(cz[4]) / max(int(cz[5]), 1) / 1e3:.1f} us each (build {int(cz[7]) / max(int(cz[5]), 1) / 1e3:.1f} us, wait for the entry {int(cz[6]) / max(int(cz[5]), 1) / 1e3:.1f} us)")
    t_edge = float(np.median([a.elapsed_time(c) for a, c in ev]))
    sl = slice(lo, lo + cnt)
    st_r, da_r = stream[sl], da[sl]
    light = da_r <= 128
    line = (f"rank {r}: edges {cnt} edge kernels {t_edge:.3f} ms | stream total {st_r.sum() / 1e6:.1f} M "
            f"(light class {st_r[light].sum() / 1e6:.1f} M, group class {st_r[~light].sum() / 1e6:.1f} M; "
            f"coop > 16384: {(st_r > 16384).sum()} edges {st_r[st_r > 16384].sum() / 1e6:.1f} M; split > 98304: {(st_r > 98304).sum()})")
    if trace:
        nct = torch.cuda.get_device_properties(0).multi_processor_count * 4
        buf = np.zeros(nct * 6, dtype=np.uint64)
        lib.dcr_paper_trace_read.restype = C.c_int
        lib.dcr_paper_trace_read(C.c_void_p(buf.ctypes.data), C.c_int(nct))
        t = buf.reshape(nct, 6).astype(np.int64)
        t0 = t[:, 0].min()
        us = lambda x: (x - t0) / 1e3
        line += (f"\n    group kernel: CTA starts max {us(t[:, 0].max()):.0f} us; coop phase ends median {us(np.median(t[:, 1])):.0f} max "
                 f"{us(t[:, 1].max()):.0f}; CTA ends p10 {us(np.quantile(t[:, 2], .1)):.0f} median {us(np.median(t[:, 2])):.0f} "
                 f"p90 {us(np.quantile(t[:, 2], .9)):.0f} max {us(t[:, 2].max()):.0f} us; longest items "
                 f"(us, phase, #deferred edges of the run) "
                 f"{[(int(t[c, 3] / 1e3), int(t[c, 4] >> 30), int((t[c, 4] >> 20) & 1023) if (t[c, 4] >> 30) == 2 else 0) for c in np.argsort(-t[:, 3])[:8]]}")
    print(line, flush=True)

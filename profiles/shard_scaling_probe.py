"""Time one rank's share of the arxiv-shaped pass for several world sizes on ONE GPU (no collectives): shows how the
per-rank kernel time scales when the shard shrinks — the fixed costs and tails that cap multi-GPU strong scaling.

    python profiles/shard_scaling_probe.py [--ncu-world W]
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "discrete-curvature-rewiring_b200"))
sys.path.insert(0, REPO)
from dcr import bfc, graph  # noqa: E402
from dcr.synth import named_graph  # noqa: E402

ei, n = named_graph("arxiv")
rowptr, col = graph.undirected_csr(ei, n)
csr = bfc.DeviceCSR.from_host(rowptr, col)
esrc, _, _ = csr.undirected_edges()
E = int(esrc.numel())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
only = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[1] == "--ncu-world" else None
worlds = [int(w) for w in os.environ.get("PROBE_WORLDS", "1,2,4,8,16").split(",")]
for world in ([only] if only else worlds):
    count = bfc.shard_count(E, 0, world)
    ws = bfc.PaperWorkspace(csr, count, chunk=max(1, (E + world - 1) // world))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
    st = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
    for a, b in ev:
        a.record(); b.record()
    for k in range(3):
        bfc.paper_flavour(csr, rank=0, world=world, ws=ws)
    for k in range(8):
        flush.zero_()
        st[k][0].record()
        bfc.paper_flavour(csr, rank=0, world=world, ws=ws, events=ev[k])
        st[k][1].record()
    torch.cuda.synchronize()
    edge = np.median([a.elapsed_time(b) for a, b in ev])
    step = np.median([a.elapsed_time(b) for a, b in st])
    print(f"world {world:2d}: rank-0 edges {count:8d}  step {step:.3f} ms  edge kernels {edge:.3f} ms  "
          f"(ideal {5.05 / world:.3f})", flush=True)

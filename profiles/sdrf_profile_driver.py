"""Times the persistent SDRF kernel alone (cora-shaped graph): python profiles/sdrf_profile_driver.py [loops]"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "discrete-curvature-rewiring_b200"))
from dcr import graph, sdrf  # noqa: E402
from dcr.synth import named_graph  # noqa: E402

name = os.environ.get("SDRF_GRAPH", "cora")
loops = int(sys.argv[1]) if len(sys.argv) > 1 else 300
ei, n = named_graph(name)
uni = np.random.RandomState(3).random_sample(loops)
rp, od = graph.networkx_order(ei, n)
best = None
for rep in range(3):
    st = sdrf.SdrfState(rp, od, max_additions=loops)
    u = torch.from_numpy(uni).cuda()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    res, log = st.run(loops, True, 0.95, 163, u)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    best = ms if best is None else min(best, ms)
    st.close()
print(os.environ.get("DCR_LIB_PATH", "default"), name, res, "best ms", round(best, 3), "us/iter", round(1e3 * best / max(res["iterations_done"], 1), 2))

import os, sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo/discrete-curvature-rewiring_b200')
from dcr import graph, sdrf, lib as L
from dcr.synth import named_graph
for name, loops in (("cora", 1000), ("wisconsin", 136)):
    ei, n = named_graph(name)
    uni = np.random.RandomState(3).random_sample(loops)
    rp, od = graph.networkx_order(ei, n)
    st = sdrf.SdrfState(rp, od, max_additions=loops)
    lib = L.load()
    buf = (C.c_ulonglong * 16)()
    lib.dcr_sdrf_phase_cycles(buf, 1)
    res, log = st.run(loops, True, 0.95, 163, torch.from_numpy(uni).cuda())
    torch.cuda.synchronize()
    lib.dcr_sdrf_phase_cycles(buf, 1)
    names = ["argmin/max", "scoring: cells", "selection", "insert rows", "supports+dirty(add)", "removal", "refresh",
             "scoring: prepare"]
    tot = sum(buf[:8])
    print(name, res["iterations_done"], "iters; cycles/iter", tot // max(res["iterations_done"], 1))
    for i, nm in enumerate(names):
        print(f"   {nm:22s} {100*buf[i]/tot:5.1f}%  {buf[i]//res['iterations_done']:8d} cyc/iter")
    ncand = log[:, 2].float().mean().item()
    print("   mean candidates", ncand, "removed", int((log[:,6]>=0).sum()))

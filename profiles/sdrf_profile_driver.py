import sys, numpy as np, torch
sys.path.insert(0,'/root/repo/discrete-curvature-rewiring_b200'); sys.path.insert(0,'/root/repo')
from dcr import graph, sdrf
from dcr.synth import named_graph
ei,n=named_graph('cora'); loops=300
uni=np.random.RandomState(3).random_sample(loops)
rp,od=graph.networkx_order(ei,n)
st=sdrf.SdrfState(rp,od,max_additions=loops)
u=torch.from_numpy(uni).cuda()
res,log=st.run(loops,True,0.95,163,u)
torch.cuda.synchronize(); print(res)

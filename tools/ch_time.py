"""Wall time of the full compute_transport_map drop-in at ChickenHeart shapes (reference 1.48 s / 3.97 s)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ot_dense
from spadot_b200 import ot_solvers
cfg = dict(ot_dense.DEFAULT_OT_CONFIG)
for (n,m) in [(747,1966),(1966,1916),(1916,1967)]:
    a,b,_,_ = ot_dense.synthetic_embeddings(n,m,20,seed=n)
    ot_solvers.compute_transport_map(a,b,dict(cfg))
    torch.cuda.synchronize(); t0=time.perf_counter()
    g = ot_solvers.compute_transport_map(a,b,dict(cfg))
    torch.cuda.synchronize(); t1=time.perf_counter()
    print(n,m,"compute_transport_map wall", t1-t0)
    # breakdown
    from spadot_b200.cuda_ops import CudaOps
    from spadot_b200 import sinkhorn
    t0=time.perf_counter(); ops=CudaOps(a,b); torch.cuda.synchronize(); t1=time.perf_counter()
    med=sinkhorn.median_cost(ops); torch.cuda.synchronize(); t2=time.perf_counter()
    ops.set_median(med); info={}
    l0=ops.launches
    st,eps=sinkhorn.solve_duality_gap(ops,np.ones(n),info=info,**cfg); torch.cuda.synchronize(); t3=time.perf_counter()
    print("  prep",t1-t0,"median",t2-t1,"solve",t3-t2,"iters",info['total_iters'],"launches",ops.launches-l0)

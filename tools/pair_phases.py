"""Phase times of analyze.transport_between for one MOSTA-shaped pair (mosta_bench data recipe)."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from spadot_b200 import analyze, ot_solvers, sinkhorn
from spadot_b200.cuda_ops import CudaOps
SPOTS = [5913, 18408, 30124, 51365, 77369, 102519, 113350, 121767]
t = int(sys.argv[1]) if len(sys.argv) > 1 else 4
lat = []
for k in (t, t + 1):
    x, _ = bench.synth(SPOTS[k], 8, 20, seed=1993)
    lat.append(x[np.random.default_rng(k).permutation(SPOTS[k])] + 0.05 * k)
def tick(label, t0):
    torch.cuda.synchronize(); t1 = time.perf_counter(); print(f"  {label}: {1e3 * (t1 - t0):.1f} ms", flush=True); return t1
for rep in range(2):
    print("rep", rep)
    t0 = time.perf_counter()
    ops = CudaOps(lat[0], lat[1]); t0 = tick("CudaOps", t0)
    info = {}
    med = sinkhorn.median_cost(ops, sinkhorn.Dist(enabled=False), info=info); t0 = tick("median " + json.dumps(info.get("phases_s")), t0)
    G = np.ones(ops.n)
    for g in range(3):
        cp = ot_solvers.solve_coupling(lat[0], lat[1], analyze.WOT_CONFIG, G=G, median=med, ops=ops, dist=sinkhorn.Dist(enabled=False))
        t0 = tick(f"solve {g} iters={cp.info['iters_per_stage']} pred_ok={ops._pred is not None and ops._pred.ok}", t0)
        G = cp.row_mass().cpu().numpy(); t0 = tick("row_mass", t0)
    cp2, growth = analyze.transport_between(lat[0], lat[1]); t0 = tick("transport_between (whole)", t0)

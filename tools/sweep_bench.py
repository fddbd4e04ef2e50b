"""Full unbalanced-OT solves (exact median + 6-stage duality-gap solver to tolerance 1e-8) at scale:
BASELINE.json configs[4] (epsilon / lambda sweep at 250k x 250k) and configs[3] (1M x 1M).  Reports
iterations, wall time and the final duality gap per setting; one JSON line per solve.

    python tools/sweep_bench.py --n 250000 --m 250000 [--full-sweep | --grid]
    torchrun --nproc-per-node 8 tools/sweep_bench.py --rows 1000000 --cols 1000000     # rows partitioned over ranks

Times are host wall-clock around synchronised regions, max over ranks.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from spadot_b200 import ot_solvers, sinkhorn  # noqa: E402
from spadot_b200.cuda_ops import CudaOps  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", "--rows", dest="n", type=int, default=250_000)   # torchrun's parser chokes on a bare --n
    ap.add_argument("--m", "--cols", dest="m", type=int, default=250_000)
    ap.add_argument("--d", type=int, default=32)
    ap.add_argument("--full-sweep", action="store_true")
    ap.add_argument("--grid", action="store_true", help="SURVEY 8d grid: eps {0.01,0.02,0.05,0.1} x 4 lambda pairs")
    ap.add_argument("--max-iter", type=float, default=3000)
    ap.add_argument("--repeat", type=int, default=1, help="solve every setting this many times (first-use effects show as a slower first line)")
    ap.add_argument("--stages", action="store_true", help="print the per-stage times (the reference's profiling switch)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        import torch.distributed as td
        td.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    dist = sinkhorn.Dist(enabled=world > 1)

    def wall(t0):
        """seconds since t0 after a device sync, max over ranks"""
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        return float(dist.max_(t).item())

    def say(**kw):
        if rank == 0:
            print(json.dumps(dict(n_gpus=world, **kw)), flush=True)

    x, y = bench.synth(a.n, a.m, a.d)
    r0, r1 = (a.n * rank) // world, (a.n * (rank + 1)) // world
    x = x[r0:r1]
    # untimed warm-up on a slice (first-use costs: allocator growth, kernel attributes, the library's NCCL communicator)
    # ALL columns: the per-iteration all-reduce then has the message size of the timed solve; at 8 ranks the first collectives of
    # a new size class cost 0.1-1.5 s once per process (profiles/r2_sweep250k_x3_8gpu.jsonl: 1.73 s, then 0.433 s, 0.433 s)
    nw = max(4096, min(a.n, 32768))
    xw, yw = x[:max(1, nw // world)], y
    ot_solvers.solve_coupling(xw, yw, dict(ot_solvers.default_config, epsilon=0.05, lambda1=0.1, lambda2=5.0, tau=1000.0),
                              ops=CudaOps(xw, yw), dist=dist, median=160.0)
    if world > 1:
        td.barrier()
    t0 = time.perf_counter()
    ops = CudaOps(x, y)
    t_prep = wall(t0)
    info = {}
    t0 = time.perf_counter()
    med = sinkhorn.median_cost(ops, dist, info=info)
    t_med = wall(t0)
    say(stage="prep+median", n=a.n, m=a.m, d=a.d, prep_s=t_prep, median_s=t_med, median=med, **info)
    settings = [(0.05, 0.1, 5.0)]
    if a.full_sweep:
        settings += [(0.05, 1.0, 50.0), (0.01, 0.1, 5.0), (0.1, 1.0, 1.0)]
    if a.grid:
        settings = [(e, l1, l2) for e in (0.01, 0.02, 0.05, 0.1) for l1, l2 in ((0.1, 5.0), (1.0, 1.0), (1.0, 50.0), (50.0, 50.0))]
    for eps, l1, l2 in [s_ for s_ in settings for _ in range(a.repeat)]:
        cfg = dict(ot_solvers.default_config, epsilon=eps, lambda1=l1, lambda2=l2, tau=1000.0, max_iter=a.max_iter, profiling=a.stages)
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        t0 = time.perf_counter()
        c0 = dist.collectives
        cp = ot_solvers.solve_coupling(x, y, cfg, median=med, ops=ops, dist=dist)
        ncoll = dist.collectives - c0
        mass = float(dist.sum_(cp.row_mass().sum().reshape(1)).item())
        dt = wall(t0)
        it = cp.info["total_iters"]
        # checksums of the potentials: every world size solves the same problem, so these must agree across N
        f_sum = float(dist.sum_(cp.f.sum().reshape(1).clone()).item())
        say(stage="solve", n=a.n, m=a.m, epsilon=eps, lambda1=l1, lambda2=l2, seconds=dt,
            iters_per_stage=cp.info["iters_per_stage"], total_iters=it, gap=cp.info["gap"],
            iters_per_sec=it / dt, plan_mass=mass, tc=ops.use_tc, converged=bool(cp.info["gap"] <= cfg["tolerance"]),
            collectives_per_iter=ncoll / max(it, 1), f_checksum=f_sum, g_checksum=float(cp.g.sum().item()),
            predicted=bool(ops._pred is not None and ops._pred.ok), native_comm=getattr(dist, "native_comm_kind", None))
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()

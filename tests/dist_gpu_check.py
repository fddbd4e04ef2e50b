"""Run under torchrun on >= 2 GPUs: the row-partitioned solve must reproduce the single-GPU solve.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ot_dense  # noqa: E402
from spadot_b200 import ot_solvers, sinkhorn  # noqa: E402
from spadot_b200.cuda_ops import CudaOps  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = dict(ot_dense.DEFAULT_OT_CONFIG)
    for (n, m, d, tc) in [(3001, 2500, 20, "off"), (5000, 4100, 32, "on")]:
        a, b, la, lb = ot_dense.synthetic_embeddings(n, m, d, seed=n)
        G = np.exp(np.random.default_rng(0).normal(0, 0.3, n))
        r0, r1 = (n * rank) // world, (n * (rank + 1)) // world
        dist = sinkhorn.Dist()
        ops = CudaOps(a[r0:r1], b, tc=tc)
        ops.PREDICT_MIN_PAIRS = 0          # predicted stabiliser on (tensor-core case): its flag rides on an all-reduce
        cp = ot_solvers.solve_coupling(a[r0:r1], b, cfg, G=G[r0:r1], ops=ops, dist=dist)
        tab = cp.transition_table(la[r0:r1], lb, 10, 10).cpu().numpy()
        # single-GPU run of the same problem on every rank
        single = sinkhorn.Dist(enabled=False)
        ops1 = CudaOps(a, b, tc=tc)
        cp1 = ot_solvers.solve_coupling(a, b, cfg, G=G, ops=ops1, dist=single)
        tab1 = cp1.transition_table(la, lb, 10, 10).cpu().numpy()
        assert cp.median == cp1.median, (cp.median, cp1.median)
        assert cp.info["iters_per_stage"] == cp1.info["iters_per_stage"], (cp.info, cp1.info)
        df = float((cp.f - cp1.f[r0:r1]).abs().max())
        dg = float((cp.g - cp1.g).abs().max())
        dt = float(np.abs(tab - tab1).max() / tab1.max())
        assert df < 1e-6 and dg < 1e-6 and dt < 1e-6, (df, dg, dt)
        if rank == 0:
            print(f"dist check {n}x{m} d={d} tc={tc}: world={world} |df|={df:.2e} |dg|={dg:.2e} table rel={dt:.2e} "
                  f"iters={cp.info['iters_per_stage']} collectives={dist.collectives} "
                  f"predicted={ops._pred is not None and ops._pred.ok} comm={getattr(dist, 'native_comm_kind', None)}", flush=True)
    td.destroy_process_group()


if __name__ == "__main__":
    main()

"""Per-iteration device time of the native sweep loop (sdb_sinkhorn_sweeps, no host synchronisation inside) across
problem sizes, against the SFU bound 2*N*M / (148 SMs x 16 ex2/clk x sm_max_mhz): shows where launch latency, tile
quantisation and pipeline fill — not the SFU — set the time.  One JSON line per size.

    python tools/size_scan.py [--sizes 2048,4096,...] [--d 32] [--sweeps 50]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from spadot_b200 import sinkhorn  # noqa: E402
from spadot_b200.cuda_ops import CudaOps  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="747x1966,1966x1916,4096x4096,5913x18408,8192x8192,16384x16384,18408x30124,32768x32768,"
                                       "65536x65536,131072x131072")
    ap.add_argument("--d", type=int, default=32)
    ap.add_argument("--sweeps", type=int, default=50)
    ap.add_argument("--tc", default="auto")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    peak = 148 * 16 * 1965e6
    for size in a.sizes.split(","):
        n, m = (int(v) for v in size.split("x"))
        x, y = bench.synth(n, m, a.d)
        ops = CudaOps(x, y, tc=a.tc)
        ops.set_median(2.0 * a.d * 2.5)
        st = sinkhorn._State(ops, np.ones(n), sinkhorn.Dist(enabled=False))
        eps, l1, l2 = 0.05, 0.1, 5.0
        args = (st, eps, l1 / (l1 + eps), l2 / (l2 + eps), float(np.log(1000.0)), float("-inf"))
        ops.fused_sweeps(*args, 10, False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.fused_sweeps(*args, a.sweeps, False)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / a.sweeps
        ideal = 2.0 * n * m / peak * 1e6
        print(json.dumps(dict(n=n, m=m, d=a.d, tc=bool(ops.use_tc), us_per_iter=us, sfu_bound_us=ideal, frac=ideal / us,
                              finite=bool(torch.isfinite(st.f).all().item()))), flush=True)


if __name__ == "__main__":
    main()

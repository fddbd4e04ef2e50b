"""The reference's own solver driver (ot_solvers.py:240-290: update_K_c + update_process_c per epsilon stage, dense
host matrices) timed with two libraries behind the same ctypes calls: libot_b200.so (this repo, device kernels) and
oracle/_ref/libot_ref.so (the reference's ot_func.cpp compiled unmodified, one CPU thread) as the baseline beside it.
One JSON line per size.  The dense contract makes the device library PCIe-bound: h2d / d2h bytes are reported.

    python tools/libot_bench.py [--sizes 747x1966,1966x1916,4096x4096] [--no-cpu]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from oracle import ot_dense, ref_lib  # noqa: E402   (CPU baseline + the shared driver only)
from spadot_b200 import ot_func  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="747x1966,1966x1916,4096x4096,8192x8192")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-max-pairs", type=float, default=2.0e7, help="largest N*M the single-thread reference is timed on")
    a = ap.parse_args()
    assert ot_func.lib.libot_b200_device_check() == 0, "needs a B200"
    dev = ref_lib.bind(ot_func.LIB_PATH)
    for size in a.sizes.split(","):
        n, m = (int(v) for v in size.split("x"))
        x, y = bench.synth(n, m, 20)
        C, _ = ot_dense.median_normalised_cost(x, y)
        G = np.ones(n)
        kw = dict(lambda1=0.1, lambda2=5.0, epsilon=0.05)
        ref_lib.duality_gap_solve(C[:64, :64].copy(), G[:64], L=dev, **kw)          # context creation, module load
        l0, h0, d0 = ot_func.counters()
        info = {}
        t0 = time.perf_counter()
        T = ref_lib.duality_gap_solve(C, G, info=info, L=dev, **kw)
        t_dev = time.perf_counter() - t0
        l1, h1, d1 = ot_func.counters()
        row = dict(n=n, m=m, solver="optimal_transport_duality_gap via update_k_double + update_process_double x 6 stages",
                   libot_b200_s=t_dev, gap=info["gap"], plan_mass=float(T.sum()), kernel_launches=l1 - l0,
                   h2d_gb=(h1 - h0) / 1e9, d2h_gb=(d1 - d0) / 1e9)
        if not a.no_cpu and ref_lib.available() and n * m <= a.cpu_max_pairs:
            t0 = time.perf_counter()
            T_ref = ref_lib.duality_gap_solve(C, G, **kw)
            row.update(libot_ref_s=time.perf_counter() - t0, cpu_threads=1,
                       max_rel_diff=float(np.abs(T - T_ref).max() / T_ref.max()))
            row["speedup"] = row["libot_ref_s"] / t_dev
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()

"""Row-pass time of the SIMT and tensor-core forms across problem sizes (development aid for the
size thresholds in CudaOps)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from spadot_b200.cuda_ops import CudaOps

def t_pass(ops, g, reps=20):
    for _ in range(3): ops.row_lse(g, 0.05)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ops.row_lse(g, 0.05)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

for n in (1024, 2048, 4096, 8192, 16384, 32768, 65536):
    x, y = bench.synth(n, n, 32)
    out = [f"n={n}"]
    for tc, mint in (("off", 1.0), ("on", 0.5), ("on", 1.0), ("on", 2.0), ("on", 4.0)):
        CudaOps.TC_ITEM_OVERHEAD_TILES = mint          # split planner's per-item overhead, in tile-times
        ops = CudaOps(x, y, tc=tc); ops.set_median(160.0)
        g = ops.zeros(n)
        out.append(f"tc={tc}/over{mint}: {t_pass(ops, g):8.1f} us")
    print("  ".join(out), flush=True)

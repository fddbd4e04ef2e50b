"""Signed error of the streamed row LSE (tensor-core and SIMT forms) against the fp64 oracle on a sample of rows:
mean (bias) and max |error|.  A negative LSE bias shows up one-to-one as a positive relative bias of the plan."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ot_dense, ot_logdomain  # noqa: E402  (checker only)
from spadot_b200.cuda_ops import CudaOps  # noqa: E402

torch.cuda.set_device(0)
for n, m, d in [(8192, 8192, 32), (40000, 30000, 32), (8192, 8192, 20)]:
    a, b, _, _ = ot_dense.synthetic_embeddings(n, m, d, seed=n)
    rng = np.random.default_rng(0)
    rows = rng.integers(0, n, 256)
    g = rng.normal(0, 0.3, m)
    med = 2.0 * d * 2.5
    for eps in (0.05, 1.0):
        want = ot_logdomain.CostOperator(a[rows], b, median=med).row_lse(g / eps, eps)
        out = dict(n=n, m=m, d=d, eps=eps)
        for tc in ("on", "off"):
            ops = CudaOps(a, b, tc=tc)
            ops.set_median(med)
            L = ops.row_lse(ops.tensor(g), eps).cpu().numpy()[rows]
            out[f"tc_{tc}_bias"] = float(np.mean(L - want))
            out[f"tc_{tc}_maxabs"] = float(np.abs(L - want).max())
        print(json.dumps(out), flush=True)

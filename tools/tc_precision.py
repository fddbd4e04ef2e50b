"""Signed error of the tensor-core streamed LSE against the fp64 oracle across the SWEEP regularisations
(BASELINE.json configs[4]: eps 0.01-0.1): mean (bias) and max |error| of row and column passes on sampled rows,
with the truncation compensation of CudaOps off / on, and the SIMT form beside it."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ot_dense, ot_logdomain  # noqa: E402  (checker only)
from spadot_b200.cuda_ops import CudaOps  # noqa: E402

torch.cuda.set_device(0)
shapes = [(3000, 2600, 32), (8192, 8192, 32), (40000, 30000, 32), (8192, 8192, 20), (6000, 5000, 10), (5000, 6000, 48)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for n, m, d in shapes:
    a, b, _, _ = ot_dense.synthetic_embeddings(n, m, d, seed=n)
    rng = np.random.default_rng(0)
    rows = rng.integers(0, n, 192)
    cols = rng.integers(0, m, 192)
    med = 2.0 * d * 2.5 * 1.0173          # not a "round" number: the exponent scale is a generic real
    for eps in (0.1, 0.05, 0.02, 0.01):
        # potentials of a realistic magnitude: a few eps-units of spread
        g = rng.normal(0, 3.0 * eps, m)
        f = rng.normal(0, 3.0 * eps, n)
        want_r = ot_logdomain.CostOperator(a[rows], b, median=med).row_lse(g / eps, eps)
        want_c = ot_logdomain.CostOperator(a, b[cols], median=med).col_lse(f / eps, eps)
        out = dict(n=n, m=m, d=d, eps=eps)
        for name, tc, comp in (("tc_nocomp", "on", 0.0), ("tc", "on", 1.0), ("simt", "off", 0.0)):
            CudaOps.TC_TRUNC_COMP_PER_KSTEP = comp
            ops = CudaOps(a, b, tc=tc)
            ops.set_median(med)
            Lr = ops.row_lse(ops.tensor(g), eps).cpu().numpy()[rows]
            Lc = ops.col_lse(ops.tensor(f), eps).cpu().numpy()[cols]
            er, ec = Lr - want_r, Lc - want_c
            out[name] = dict(row_bias=float(er.mean()), row_max=float(np.abs(er).max()), col_bias=float(ec.mean()),
                             col_max=float(np.abs(ec).max()))
            del ops
        print(json.dumps(out), flush=True)

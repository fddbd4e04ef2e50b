"""Probe of tcgen05.mma kind::f16 arithmetic through the C ABI (sdb_lse_pass_tc with one live column, scale = 1,
bias = 0: the row maximum that comes back IS the fp32 accumulator value).

Questions: (1) how is a K=16 step added into the fp32 accumulator (round-to-nearest or truncation, toward zero or toward
-inf), (2) how many bits survive inside one K=16 step when large products cancel (internal alignment width).
The MMA issue order of the pass is  A_hi.B_lo,  A_lo.B_hi,  A_hi.B_hi  (sdb_tc.cu), dp = 16 => one K step per chain."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spadot_b200 import _lib  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda:0")
_lib.require_device()
DP = 16


def run(x_hi, x_lo, y_hi, y_lo):
    """x_*: (rows, 16) float; y_*: (16,) float -> accumulator value per row (float32)."""
    rows = x_hi.shape[0]
    P = np.zeros((256, 2 * DP), dtype=np.float16)
    P[:rows, :DP] = x_hi
    P[:rows, DP:] = x_lo
    Q = np.zeros((256, 2 * DP), dtype=np.float16)
    Q[0, :DP] = y_hi
    Q[0, DP:] = y_lo
    bias = np.full(256, -1.0e30, dtype=np.float32)
    bias[0] = 0.0
    p16, q16 = torch.from_numpy(P).to(dev), torch.from_numpy(Q).to(dev)
    b = torch.from_numpy(bias).to(dev)
    partial = torch.zeros((1, rows, 2), dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.call("sdb_lse_pass_tc", p16.data_ptr(), rows, 256, q16.data_ptr(), 1, 256, DP, b.data_ptr(), 1.0, 1, 148,
              partial.data_ptr(), st)
    torch.cuda.synchronize()
    return partial[0, :, 0].cpu().numpy().astype(np.float64), partial[0, :, 1].cpu().numpy()


out = {}
# ---- (1) accumulate rounding: chain 1 (A_hi.B_lo) writes +-2^24, chain 3 (A_hi.B_hi) adds a small value
y_hi = np.zeros(16); y_hi[1:] = 1.0
y_lo = np.zeros(16); y_lo[0] = 4096.0
cases = [(4096.0, [1.5]), (-4096.0, [-1.5]), (4096.0, [-0.25]), (-4096.0, [0.25]), (4096.0, [3.0]), (4096.0, [1.0]),
         (4096.0, [1.25, 1.25]), (4096.0, [0.75, 0.75]), (-4096.0, [-1.25, -1.25]), (4096.0, [1.0, 0.5, 0.25, 0.125]),
         (4096.0, [-0.5, -0.5]), (4096.0, [5.0, -3.5])]
xh = np.zeros((len(cases), 16))
for i, (big, adds) in enumerate(cases):
    xh[i, 0] = big
    xh[i, 1:1 + len(adds)] = adds
got, _ = run(xh, np.zeros_like(xh), y_hi, y_lo)
out["accumulate"] = [dict(acc=big * 4096.0, adds=adds, exact=big * 4096.0 + sum(adds), got=float(v), got_minus_acc=float(v - big * 4096.0))
                     for (big, adds), v in zip(cases, got)]
# ---- (2) internal width of one K step: big - big + small inside chain 3 only (accumulator starts ~0)
y_hi = np.ones(16); y_lo = np.zeros(16)
res = []
for pos_neg in (2, 3, 9, 15):
    for pos_small in (1, 7, 14):
        if pos_small == pos_neg:
            continue
        ks = list(range(0, 25))
        xh = np.zeros((len(ks), 16))
        for i, k in enumerate(ks):
            xh[i, 0] = 32768.0
            xh[i, pos_neg] = -32768.0
            xh[i, pos_small] = 2.0 ** -k
        got, _ = run(xh, np.zeros_like(xh), y_hi, y_lo)
        kept = [k for k, v in zip(ks, got) if v == 2.0 ** -k]
        part = [(k, float(v)) for k, v in zip(ks, got) if v != 2.0 ** -k and v != 0.0]
        res.append(dict(pos_neg=pos_neg, pos_small=pos_small, smallest_kept_exp=-max(kept) if kept else None,
                        bits_below_big=15 + max(kept) if kept else None, partial=part[:4]))
out["cancel_within_step"] = res
# ---- (3) products of different magnitude in one step without cancellation: 2^15*2^15 + 2^-k (true sum needs 30+k bits)
res = []
y_hi = np.ones(16); y_hi[0] = 32768.0
for k in range(0, 12):
    xh = np.zeros((1, 16)); xh[0, 0] = 32768.0; xh[0, 1] = 2.0 ** (7 - k)      # 2^30 + 2^(7-k): fp32 ulp at 2^30 is 2^7
    got, _ = run(xh, np.zeros_like(xh), y_hi, y_lo)
    res.append(dict(small=2.0 ** (7 - k), got_minus_big=float(got[0] - 2.0 ** 30)))
out["round_of_small_addend"] = res
# ---- (4) sticky/half cases inside one step: 2^30 + 2^6 (half ulp) + 2^6 (another product)
xh = np.zeros((4, 16)); y_hi = np.ones(16); y_hi[0] = 32768.0
xh[:, 0] = 32768.0
xh[0, 1] = 64.0; xh[0, 2] = 64.0             # two half-ulps: exact sum is representable (2^30 + 128)
xh[1, 1] = 64.0; xh[1, 2] = 32.0; xh[1, 3] = 32.0
xh[2, 1:9] = 16.0                            # eight eighth-ulps
xh[3, 1] = 96.0                              # 0.75 ulp
got, _ = run(xh, np.zeros_like(xh), y_hi, y_lo)
out["sub_ulp_addends"] = [float(v - 2.0 ** 30) for v in got]
print(json.dumps(out, indent=1))

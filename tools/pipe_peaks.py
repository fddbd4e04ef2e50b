"""Measured pipe peaks of this B200 ("to be confirmed by a micro-benchmark on the box").

Runs sdb_pipe_peak for MUFU.EX2, FFMA, FFMA2 and the epilogue's instruction mix, timed with CUDA events after
warm-up, while nvidia-smi samples the SM clock; prints one JSON object (also importable: `measure(device)`)."""
import ctypes
import json
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spadot_b200 import _lib  # noqa: E402

KINDS = {0: ("mufu_ex2", "Tex2/s", 1.0), 1: ("ffma", "TFLOP/s", 2.0), 2: ("ffma2", "TFLOP/s", 2.0), 3: ("epilogue_mix", "Tex2/s", 1.0)}


def _clock(index):
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-i", str(index)],
                             capture_output=True, text=True, timeout=10).stdout.strip().split(",")
        return float(out[0]), float(out[1])
    except Exception:
        return None, None


def measure(device=None, iters=4096, reps=5):
    _lib.require_device()
    dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    n_ctas = n_sm * 2 * 4
    out = torch.zeros(n_ctas * 512, dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    res = {"n_sm": n_sm}
    for kind, (name, unit, mult) in KINDS.items():
        ops = ctypes.c_double(0.0)
        for _ in range(2):
            _lib.call("sdb_pipe_peak", kind, n_ctas, iters, out.data_ptr(), ctypes.addressof(ops), st)
        torch.cuda.synchronize(dev)
        best, clocks = None, []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.call("sdb_pipe_peak", kind, n_ctas, iters, out.data_ptr(), ctypes.addressof(ops), st)
            e1.record()
            c, cmax = _clock(dev.index or 0)        # sampled while the kernel runs (the launch is asynchronous)
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1)
            if c:
                clocks.append(c)
            best = ms if best is None else min(best, ms)
        rate = ops.value * mult / (best * 1e-3) / 1e12
        res[name] = dict(value=rate, unit=unit, ms=best, sm_mhz_sampled=(sorted(clocks)[len(clocks) // 2] if clocks else None),
                         per_clk_per_sm_at_sampled=(rate * 1e12 / mult / n_sm / (sorted(clocks)[len(clocks) // 2] * 1e6)) if clocks else None)
        res["sm_max_mhz"] = cmax
    return res


if __name__ == "__main__":
    print(json.dumps(measure()))

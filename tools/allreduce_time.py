"""Latency of the in-library all-reduce (sdb_nccl_allreduce_sum_f64) beside torch.distributed's, per message size,
under torchrun: the collective of every row-partitioned Sinkhorn iteration is M + 2 doubles.  One JSON line per size."""
import json
import os
import sys

import torch
import torch.distributed as td

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spadot_b200 import _lib, sinkhorn  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    td.init_process_group("nccl", device_id=dev)
    dist = sinkhorn.Dist()
    comm = dist.native_comm(dev)
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    for count in (4, 18410, 30126, 51367, 77371, 121769, 250002, 1000002):
        buf = torch.ones(count, dtype=torch.float64, device=dev)
        res = {"n_gpus": world, "doubles": count}
        for name, call in (("native", lambda: _lib.check(lib.sdb_nccl_allreduce_sum_f64(comm, buf.data_ptr(), count, stream), "ar")),
                           ("torch", lambda: td.all_reduce(buf)),
                           ("torch_max", lambda: td.all_reduce(buf, op=td.ReduceOp.MAX))):
            if name == "native" and comm is None:
                continue
            first = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            first[0].record(); call(); first[1].record(); torch.cuda.synchronize()
            for _ in range(5):
                call()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            td.barrier(); torch.cuda.synchronize()
            e0.record()
            for _ in range(50):
                call()
            e1.record(); torch.cuda.synchronize()
            res[name + "_first_ms"] = first[0].elapsed_time(first[1])
            res[name + "_us"] = 1e3 * e0.elapsed_time(e1) / 50
            buf.fill_(1.0)
        if rank == 0:
            print(json.dumps(res), flush=True)
    td.destroy_process_group()


if __name__ == "__main__":
    main()

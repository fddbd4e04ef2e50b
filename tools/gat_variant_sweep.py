"""A/B of the tile kernels' load-pipelining knobs (SDB_GAT_AGG_VARIANT / SDB_GAT_BDST_VARIANT, csrc/sdb_gat.cu) on the
SYN-T-shaped graph of tools/gat_kernel_bench.py: forward = aggregation kernel, backward = by-destination + by-source."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spadot_b200 import gat, graph


def timed(fn_f, go, reps=6):
    for _ in range(2):
        o = fn_f(); o.backward(go)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(reps):
        e[0].record(); o = fn_f(); e[1].record(); o.backward(go); e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    return tf / reps, tb / reps


def main():
    dev = torch.device("cuda:0")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    H, C, k = 4, 512, 30
    coords = np.random.default_rng(0).uniform(0, 1000, size=(n, 2))
    ei = graph.spatial_edge_index(coords, k)
    g = gat.graph_for(ei, n, True, torch.from_numpy(coords).to(dev))
    for dtype in (torch.float32, torch.float64):
        feat = torch.randn(n, H, C, dtype=dtype, device=dev, requires_grad=True)
        a_s = torch.randn(n, H, dtype=dtype, device=dev, requires_grad=True)
        a_d = torch.randn(n, H, dtype=dtype, device=dev, requires_grad=True)
        go = torch.randn(n, H, C, dtype=dtype, device=dev)
        f = lambda: gat._EdgeSoftmaxAggregate.apply(feat, a_s, a_d, g, 0.2)
        for av in range(6):
            os.environ["SDB_GAT_AGG_VARIANT"] = str(av); os.environ["SDB_GAT_BDST_VARIANT"] = "0"
            tf, tb = timed(f, go)
            print(json.dumps(dict(dtype=str(dtype), agg_variant=av, bdst_variant=0, fwd_ms=round(tf, 4), bwd_ms=round(tb, 4))), flush=True)
        for bv in range(1, 5):
            os.environ["SDB_GAT_AGG_VARIANT"] = "0"; os.environ["SDB_GAT_BDST_VARIANT"] = str(bv)
            tf, tb = timed(f, go)
            print(json.dumps(dict(dtype=str(dtype), agg_variant=0, bdst_variant=bv, fwd_ms=round(tf, 4), bwd_ms=round(tb, 4))), flush=True)
        del feat, a_s, a_d, go


if __name__ == "__main__":
    main()

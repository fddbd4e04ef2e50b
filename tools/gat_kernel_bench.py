"""K2/K2b kernel timing against the HBM roofline on a SYN-T-shaped graph:
n nodes, k=30 neighbours + self loop, H=4 heads x C=512 channels.  Algorithmic bytes:
forward  E*H*C*s (row gathers) + n*H*C*s (write) + E*(4 + 2*H*s)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spadot_b200 import gat, graph

def main():
    dev = torch.device("cuda:0")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    H, C, k = 4, 512, 30
    coords = np.random.default_rng(0).uniform(0, 1000, size=(n, 2))
    ei = graph.spatial_edge_index(coords, k)
    g = gat.graph_for(ei, n, True, torch.from_numpy(coords).to(dev))      # CTA order: Z-curve of the coordinates, as in training
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    out = {}
    modes = (("per_node", "0"), ("tiles", "1"))          # SDB_GAT_TILES: the per-node kernels vs the tile kernels (default)
    for (mode, flag), (dtype, s) in ((m, d) for m in modes for d in ((torch.float32, 4), (torch.float64, 8))):
        os.environ["SDB_GAT_TILES"] = flag
        feat = torch.randn(n, H, C, dtype=dtype, device=dev, requires_grad=True)
        a_s = torch.randn(n, H, dtype=dtype, device=dev, requires_grad=True)
        a_d = torch.randn(n, H, dtype=dtype, device=dev, requires_grad=True)
        go = torch.randn(n, H, C, dtype=dtype, device=dev)
        def fwd():
            return gat._EdgeSoftmaxAggregate.apply(feat, a_s, a_d, g, 0.2)
        for _ in range(3):
            o = fwd(); o.backward(go)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        reps = 5
        tf = tb = 0.0
        for _ in range(reps):
            e[0].record(); o = fwd(); e[1].record(); o.backward(go); e[2].record()
            torch.cuda.synchronize()
            tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
        tf, tb = tf / reps, tb / reps
        E = g.E
        # (a) traffic if every per-edge row gather went to memory (what an unordered launch pays, served by L2 here)
        gather_f = E * H * C * s + n * H * C * s + E * (4 + 2 * H * s)
        gather_b = 2 * E * H * C * s + 2 * n * H * C * s + E * (8 + 4 * H * s)
        # (b) SURVEY.md §8d algorithmic HBM bytes: every feature row read once, outputs written once
        hbm_f = 2 * n * H * C * s + E * (4 + 2 * H * s)
        hbm_b = 3 * n * H * C * s + E * (8 + 4 * H * s)
        out[mode + "_" + str(dtype).split(".")[1]] = dict(fwd_ms=tf, bwd_ms=tb, fwd_gather_GBs=gather_f / tf / 1e6, bwd_gather_GBs=gather_b / tb / 1e6,
                                             fwd_hbm_algorithmic_GBs=hbm_f / tf / 1e6, bwd_hbm_algorithmic_GBs=hbm_b / tb / 1e6,
                                             fwd_frac_of_hbm=hbm_f / tf / 1e6 / hbm, bwd_frac_of_hbm=hbm_b / tb / 1e6 / hbm)
    os.environ.pop("SDB_GAT_TILES", None)
    print(json.dumps(dict(n=n, edges=g.E, H=H, C=C, hbm_peak_gbs=hbm, **out)))

if __name__ == "__main__":
    main()

"""BASELINE.json configs[2]: synthetic timepoint of 100k spots x 2000 SVGs, latent dim 32 (16 GP + 16 GAT dims),
150 inducing points, k=30 spatial graph — SVGP+GAT training throughput of this repo's model on one GPU, float64.

The reference cannot run this shape at all: _Cal_Spatial_Net builds a dense N x N adjacency (80 GB at 100k,
utils/_utils.py:98-100), kernel_matrix(diag_only=True) builds an N x N block (model/svgp.py:43-45) and PyG's
GATConv materialises (E,H,C) messages (3.1M x 4 x 512 fp64 = 51 GB per layer).  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spadot_b200 import graph, model as product  # noqa: E402


def run(argv=None):
    """One measurement; returns the result dict (bench.py's aux_syn_t_step calls this with a short batch list)."""
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--genes", type=int, default=2000)
    ap.add_argument("--z", type=int, default=32)
    ap.add_argument("--inducing", type=int, default=150)
    ap.add_argument("--batches", type=int, default=4)
    ap.add_argument("--name", default="SYN-T")
    ap.add_argument("--epoch2", action="store_true", help="also time the steps of a second epoch (batches and their graphs kept by the loader)")
    ap.add_argument("--top-kernels", type=int, default=0, help="print the N kernels with the most device time of one step to stderr")
    ap.add_argument("--host-profile", action="store_true", help="cProfile one more step and print the top host-side entries to stderr")
    a = ap.parse_args(argv)
    dev = torch.device("cuda:0")
    n, genes, z, m_ind, n_batches = a.n, a.genes, a.z, a.inducing, a.batches
    rng = np.random.default_rng(0)
    raw = rng.uniform(0, 20000, size=(n, 2))
    loc = (raw - raw.mean(0)) / raw.std(0)
    t0 = time.perf_counter()
    ei = graph.spatial_edge_index(raw, graph.knn_cutoff(n), device=dev)
    torch.cuda.synchronize()
    t_graph = time.perf_counter() - t0
    y = torch.randn(n, genes, dtype=torch.float64, device=dev).clamp_(-10, 10)
    x = torch.from_numpy(loc).to(dev)
    cfg = dict(input_dim=genes, z_dim=z, dtype=torch.float64, device=dev, svgp_encoder_layers=[256, 64], gat_encoder_hidden=512,
               gat_attention_heads=4, decoder_layers=[64, 256], kernel_type="Gaussian", kernel_scale=0.1, timepoints=["t"])
    dl = dict(inducing_points={"t": loc[rng.choice(n, m_ind, replace=False)]}, N_train={"t": n})
    net = product.SpaDOT(cfg, dl).to(dev)
    opt = torch.optim.AdamW(net.parameters(), lr=3e-4)
    batches = []
    t0 = time.perf_counter()
    for nodes, lei, ns in graph.two_hop_batches(ei, n, batch_size=512):
        batches.append((nodes, lei, ns))
        if len(batches) == n_batches + 1:
            break
    torch.cuda.synchronize()
    t_sample = (time.perf_counter() - t0) / len(batches)

    def step(b):
        nodes, lei, ns = b
        recon, skl, gkl, align, _ = net.forward(x[nodes], y[nodes], lei, "t", ns)
        loss = 0.1 * recon - 0.5 * skl + 1e-4 * gkl + 0.1 * align
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 0.3)
        opt.step()
        return float(loss)

    step(batches[0])                                   # warm-up (cuBLAS/cuSOLVER handles, graph caches)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    losses, step_times = [], []
    for b in batches[1:]:
        t1 = time.perf_counter()
        losses.append(step(b))                       # float(loss) synchronises
        step_times.append(time.perf_counter() - t1)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n_batches
    # from the second epoch on: the loader (graph.TwoHopBatches) hands out the batches it sampled in the first epoch together
    # with their prepared graphs (CSR both ways, CTA order, prefix plans) - the reference's NeighborLoader is not shuffled
    epoch2 = None
    if a.epoch2:
        from spadot_b200 import gat
        prepared = [(nodes, gat.CsrGraph(lei, int(nodes.numel()), True, x[nodes]), ns) for nodes, lei, ns in batches[1:]]
        for b in prepared:
            step(b)                                   # first epoch over these batches: fills the graphs' plan / order caches
        torch.cuda.synchronize()
        times2 = []
        for b in prepared:
            t1 = time.perf_counter()
            step(b)
            times2.append(time.perf_counter() - t1)
        epoch2 = dict(step_times_s=[round(t, 4) for t in times2], train_step_s=sum(times2) / len(times2),
                      graph_bytes_per_batch=int(np.mean([b[1].nbytes() for b in prepared])))
    # where the step goes: device time by kernel family over one more step (torch.profiler, CUDA activities)
    breakdown = None
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step(batches[1])
            torch.cuda.synchronize()
        fam = {}
        for ev in prof.key_averages():
            name, t = ev.key, getattr(ev, "device_time_total", None) or getattr(ev, "cuda_time_total", 0.0)
            low = name.lower()
            if "gat_" in low:
                k = "gat edge-softmax / aggregation kernels (K2/K2b)"
            elif any(w in low for w in ("gemm", "cutlass", "xmma", "cublas", "gemv")):
                k = "dense GEMMs (cuBLAS fp64: GAT lin layers, MLPs, SVGP algebra)"
            elif any(w in low for w in ("kernel_block", "quad_form", "kernel_diag")):
                k = "SVGP kernel blocks (K1)"
            elif any(w in low for w in ("cusolver", "potrf", "trsm", "getrf", "syrk", "laswp", "inverse")):
                k = "factorisations (cuSOLVER / MAGMA)"
            elif "memcpy" in low or "memset" in low:
                k = "copies / memsets"
            else:
                k = "elementwise / reductions / indexing (torch)"
            fam[k] = fam.get(k, 0.0) + float(t)
        if a.top_kernels:
            top = sorted(((float(getattr(ev, "device_time_total", None) or getattr(ev, "cuda_time_total", 0.0)), ev.count, ev.key)
                          for ev in prof.key_averages()), reverse=True)[:a.top_kernels]
            for t, cnt, name in top:
                print(f"{t / 1e3:8.3f} ms  x{cnt:<4d} {name[:150]}", file=sys.stderr)
        tot = sum(fam.values())
        breakdown = {k: dict(ms=v / 1e3, share=v / tot) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])}
    except Exception as exc:
        breakdown = dict(error=repr(exc)[:200])
    if a.host_profile:
        import cProfile
        import pstats
        from spadot_b200 import gat
        gat._GRAPH_CACHE.clear()                       # like a fresh batch of the loader: the sub-graph is built inside the step
        pr = cProfile.Profile()
        pr.enable()
        step(batches[2 % len(batches)])
        torch.cuda.synchronize()
        pr.disable()
        pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(45)
    sub_nodes = int(np.mean([b[0].numel() for b in batches[1:]]))
    sub_edges = int(np.mean([b[1].shape[1] for b in batches[1:]]))
    # full-timepoint inference (all_latent_samples) without n x n temporaries
    net.eval()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lat = net.all_latent_samples(loc, y, ei, "t")
    torch.cuda.synchronize()
    t_inf = time.perf_counter() - t0
    return dict(workload=f"{a.name} one timepoint: {n} spots x {genes} genes, z={z}, {m_ind} inducing, k=30, fp64",
                          graph_build_s=t_graph, sample_batch_s=t_sample, subgraph_nodes=sub_nodes, subgraph_edges=sub_edges,
                          train_step_s=dt, step_times_s=[round(t, 4) for t in step_times], seeds_per_s=512 / dt, losses=losses, all_latent_samples_s=t_inf,
                          latent_shape=list(lat.shape), peak_mem_gb=torch.cuda.max_memory_allocated() / 1e9,
                          step_breakdown_device_time=breakdown, second_epoch=epoch2)


def main(argv=None):
    print(json.dumps(run(argv)))


if __name__ == "__main__":
    main()

"""`train` inner loop at ChickenHeart shapes on one GPU (the reference logs 2-3 s per epoch
for its 14 mini-batches): 4 timepoints (747/1966/1916/1967 spots), 2954 genes, z=20, 135/354/389/322
inducing points, kNN graph + 2-hop batches of 512 seeds, float64, AdamW(3e-4), grad-clip 0.3
(SpaDOT/utils/_train_utils.py:155-236 without the k-means / OT regularisers, which are host-side).

`reference` = oracle/model_ref.py (the reference's formulas in plain torch on the SAME GPU);
`spadot_b200` = spadot_b200/model.py.  Same initial weights, same reparameterisation noise; prints one JSON line
with seconds per epoch for both and the largest relative difference between the two loss traces.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model_ref  # noqa: E402
from spadot_b200 import graph, model as product  # noqa: E402

SIZES = {"D4": 747, "D7": 1966, "D10": 1916, "D14": 1967}
INDUCING = {"D4": 135, "D7": 354, "D10": 389, "D14": 322}
GENES = 2954


def make_data(dev, seed=1993):
    rng = np.random.default_rng(seed)
    data = {}
    for tp, n in SIZES.items():
        r = np.sqrt(rng.uniform(0, 1, n)) * 3000
        th = rng.uniform(0, 2 * np.pi, n)
        raw = np.stack([r * np.cos(th), r * np.sin(th)], 1) + 4000
        loc = (raw - raw.mean(0)) / raw.std(0)                       # _obtain_tp_loc_info: StandardScaler per tp
        y = np.clip(rng.normal(size=(n, GENES)), -10, 10)
        ei = graph.spatial_edge_index(raw, graph.knn_cutoff(n), device=dev)    # kNN on RAW coordinates (_utils.py:65)
        batches = [(nodes, lei, ns) for nodes, lei, ns in graph.two_hop_batches(ei, n, batch_size=512)]
        data[tp] = dict(loc=torch.from_numpy(loc).to(dev), y=torch.from_numpy(y).to(dev), batches=batches,
                        inducing=loc[rng.choice(n, INDUCING[tp], replace=False)], n=n)
    return data


def build(cls, data, dev, seed):
    cfg = dict(input_dim=GENES, z_dim=20, dtype=torch.float64, device=dev, svgp_encoder_layers=[256, 64],
               gat_encoder_hidden=512, gat_attention_heads=4, decoder_layers=[64, 256], kernel_type="Gaussian",
               kernel_scale=0.1, timepoints=list(SIZES))
    dl = dict(inducing_points={tp: d["inducing"] for tp, d in data.items()}, N_train={tp: d["n"] for tp, d in data.items()})
    gen = torch.Generator(device=dev).manual_seed(seed)
    return cls(cfg, dl, noise_fn=lambda t: torch.randn(t.shape, dtype=t.dtype, device=t.device, generator=gen)).to(dev), gen


def epoch(model, opt, data, beta1=0.5):
    trace = []
    for tp, d in data.items():
        for nodes, lei, ns in d["batches"]:
            recon, skl, gkl, align, _ = model.forward(d["loc"][nodes], d["y"][nodes], lei, tp, ns)
            elbo = 0.1 * recon - beta1 * skl + 1e-4 * gkl + 0.1 * align          # config.yaml lambda1/beta2/omiga1
            opt.zero_grad()
            elbo.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 0.3)
            opt.step()
            trace.append(elbo.detach())
    return torch.stack(trace).cpu().numpy()


def run(dev=None, with_reference=True, with_cpu=True):
    dev = dev or torch.device("cuda:0")
    torch.manual_seed(0)
    data = make_data(dev)
    ref, gen_r = build(model_ref.SpaDOTRef, data, dev, 7)
    mine, gen_m = build(product.SpaDOT, data, dev, 7)
    mine.load_state_dict(ref.state_dict(), strict=False)
    out = {}
    traces = {}
    arms = (("reference", ref, gen_r), ("spadot_b200", mine, gen_m)) if with_reference else (("spadot_b200", mine, gen_m),)
    for name, model, gen in arms:
        opt = torch.optim.AdamW(model.parameters(), lr=3e-4)
        gen.manual_seed(7)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tr1 = epoch(model, opt, data)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        tr2 = epoch(model, opt, data)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        out[name] = dict(first_epoch_s=t1 - t0, second_epoch_s=t2 - t1)
        traces[name] = np.concatenate([tr1, tr2])
    n_batches = len(traces["spadot_b200"]) // 2
    res = dict(workload=f"ChickenHeart-shaped synthetic (4 timepoints, 6596 spots, 2954 genes, z=20), {n_batches} batches/epoch, fp64, "
                        "fwd+bwd+clip+AdamW", **out)
    if with_reference:
        rel = np.abs(traces["reference"] - traces["spadot_b200"]) / np.abs(traces["reference"])
        res.update(speedup_second_epoch=out["reference"]["second_epoch_s"] / out["spadot_b200"]["second_epoch_s"],
                   loss_trace_max_rel_diff=float(rel.max()))
    res.update(loss_first=float(traces["spadot_b200"][0]), loss_last=float(traces["spadot_b200"][-1]))
    if with_cpu:
        res["host_cpu_sample"] = cpu_sample(data, mine, dev)
    return res


def cpu_sample(data, mine, dev, tp="D4"):
    """BASELINE's "train s/epoch ... vs host CPU", on a bounded sample: ONE optimiser step (fwd + bwd + clip + AdamW) of
    the first mini-batch of the smallest timepoint, reference formulas in torch on the host cores vs this repo's modules
    on the GPU, same weights."""
    d = data[tp]
    nodes, lei, ns = d["batches"][0]
    cpu = torch.device("cpu")
    cfg = dict(input_dim=GENES, z_dim=20, dtype=torch.float64, device=cpu, svgp_encoder_layers=[256, 64], gat_encoder_hidden=512,
               gat_attention_heads=4, decoder_layers=[64, 256], kernel_type="Gaussian", kernel_scale=0.1, timepoints=[tp])
    ref = model_ref.SpaDOTRef(cfg, dict(inducing_points={tp: d["inducing"]}, N_train={tp: d["n"]}))
    ref.load_state_dict({k: v.cpu() for k, v in mine.state_dict().items() if k in ref.state_dict()}, strict=False)
    x_c, y_c, e_c = d["loc"][nodes].cpu(), d["y"][nodes].cpu(), lei.cpu()

    def one_step(model, x, y, e):
        opt = torch.optim.AdamW(model.parameters(), lr=3e-4)
        recon, skl, gkl, align, _ = model.forward(x, y, e, tp, ns)
        loss = 0.1 * recon - 0.5 * skl + 1e-4 * gkl + 0.1 * align
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 0.3)
        opt.step()
        return float(loss)

    t0 = time.perf_counter()
    one_step(ref, x_c, y_c, e_c)
    t_cpu = time.perf_counter() - t0
    one_step(mine, d["loc"][nodes], d["y"][nodes], lei)         # warm
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    one_step(mine, d["loc"][nodes], d["y"][nodes], lei)
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    return dict(sample=f"one optimiser step, timepoint {tp} ({d['n']} spots, {len(d['inducing'])} inducing points), {ns} seeds, "
                       f"sub-graph of {int(nodes.numel())} nodes, fp64", reference_cpu_s=t_cpu, cpu_threads=torch.get_num_threads(),
                spadot_b200_gpu_s=t_gpu, ratio=t_cpu / t_gpu)


if __name__ == "__main__":
    print(json.dumps(run()))

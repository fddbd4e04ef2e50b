"""Train inner-loop pieces at ChickenHeart shapes, on one GPU:

  * SVGP: z/2 = 10 latent dims x (approximate_posterior_params + variational_loss), b=512, m=389, fwd+bwd
  * GAT encoder: 3 x GATConv (2954 -> 4x512 -> 4x512 -> 512) + Linear on a 2-hop sub-graph, fwd+bwd, float64

`reference` = the oracle restatements of SpaDOT/model/svgp.py and PyG GATConv run with plain torch ops on the
SAME GPU (what `SpaDOT train --device cuda:0` executes); `spadot_b200` = this repo's modules.  Prints one JSON line.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gat_ref, svgp_ref  # noqa: E402
from spadot_b200 import gat, graph, svgp  # noqa: E402


def timeit(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    b, m, N, zh = 512, 389, 1916, 10
    g = torch.Generator().manual_seed(0)
    z = (torch.rand(m, 2, generator=g, dtype=torch.float64) * 4 - 2)
    x = (torch.rand(b, 2, generator=g, dtype=torch.float64) * 4 - 2).to(dev)
    Y = torch.randn(b, zh, generator=g, dtype=torch.float64).to(dev).requires_grad_(True)
    NZ = (torch.rand(b, zh, generator=g, dtype=torch.float64) + 0.5).to(dev).requires_grad_(True)
    ref = svgp_ref.SVGPRef(z.numpy(), N).to(dev)
    cfg = dict(dtype=torch.float64, device=dev, kernel_type="Gaussian", kernel_scale=0.1)
    mine = svgp.SVGP(cfg, z.numpy(), N_train=N)

    def svgp_step(model):
        tot = 0
        for l in range(zh):                                   # SpaDOT.py:57-66
            mean, B, mu_hat, A_hat = model.approximate_posterior_params(x, x, Y[:, l], NZ[:, l])
            l3, kl = model.variational_loss(x, Y[:, l], NZ[:, l], mu_hat, A_hat)
            tot = tot + l3 + kl
        Y.grad = NZ.grad = None
        tot.backward()
        return float(tot)

    def svgp_step_batched():
        mean, var, L3, KL = mine.posterior_and_loss_all_dims(x, Y, NZ)
        tot = L3.sum() + KL.sum()
        Y.grad = NZ.grad = None
        tot.backward()
        return float(tot)

    v_ref, v_mine, v_bat = svgp_step(ref), svgp_step(mine), svgp_step_batched()
    t_ref, t_mine = timeit(lambda: svgp_step(ref), 1, 3), timeit(lambda: svgp_step(mine), 3, 10)
    t_bat = timeit(svgp_step_batched, 3, 10)

    # GAT encoder on a 2-hop batch of a 1916-spot timepoint
    n, genes = 1916, 2954
    coords = np.random.default_rng(0).uniform(0, 6000, size=(n, 2))
    ei = graph.spatial_edge_index(coords, graph.knn_cutoff(n))
    nodes, lei, n_seeds = next(iter(graph.two_hop_batches(ei, n, batch_size=512)))
    feats = torch.randn(nodes.numel(), genes, dtype=torch.float64, device=dev)
    enc_ref = gat_ref.GATEncoderRef(genes, 10, 512, 4).double().to(dev)
    enc = gat.GATEncoder(genes, 10, 512, 4).double().to(dev)
    enc.load_state_dict(enc_ref.state_dict())

    def gat_step(model):
        model.zero_grad(set_to_none=True)
        mu, var = model(feats, lei)
        loss = (mu[:n_seeds] ** 2).sum() + var[:n_seeds].sum()
        loss.backward()
        return float(loss)

    g_ref, g_mine = gat_step(enc_ref), gat_step(enc)
    tg_ref, tg_mine = timeit(lambda: gat_step(enc_ref), 2, 5), timeit(lambda: gat_step(enc), 2, 5)
    print(json.dumps(dict(
        svgp=dict(shape=f"b={b} m={m} latent_dims={zh} fp64 fwd+bwd", reference_ms=t_ref, spadot_b200_ms=t_mine,
                  speedup=t_ref / t_mine, loss_rel_diff=abs(v_ref - v_mine) / abs(v_ref),
                  batched_all_dims_ms=t_bat, batched_speedup=t_ref / t_bat, batched_loss_rel_diff=abs(v_ref - v_bat) / abs(v_ref)),
        gat=dict(shape=f"sub-graph nodes={nodes.numel()} edges={lei.shape[1]} genes={genes} 4x512 fp64 fwd+bwd",
                 reference_ms=tg_ref, spadot_b200_ms=tg_mine, speedup=tg_ref / tg_mine, loss_rel_diff=abs(g_ref - g_mine) / abs(g_ref)))))


if __name__ == "__main__":
    main()

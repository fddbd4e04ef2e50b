"""Generate the committed golden vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes small .npz fixtures next to this file.  The reference modules are imported by
file path (oracle/reference_loader.py); the OT solver runs through its own shipped
libot.so (C inner loop, use_C=True), i.e. exactly what `SpaDOT train` executes
(SpaDOT/utils/_train_utils.py:318) and the vendored twin of what wot runs in
`SpaDOT analyze` (SpaDOT/utils/_analyze_utils.py:124-126).
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ot_dense, reference_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CFG = dict(ot_dense.DEFAULT_OT_CONFIG)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def ot_case(ref, name, n, m, d, seed, cfg, G=None, full_plan=True, n_samples=4000):
    a, b, la, lb = ot_dense.synthetic_embeddings(n, m, d, seed=seed)
    C = ref.sklearn.metrics.pairwise.pairwise_distances(a, b, metric="sqeuclidean", n_jobs=1)
    med = float(np.median(C))
    Cn = C / med
    G0 = np.ones(n) if G is None else G
    # (1) single duality-gap solve on the median-normalised cost (ot_solvers.py:164)
    plan_gap = quiet(ref.optimal_transport_duality_gap, C=Cn.copy(), G=G0.copy(), **cfg)
    # (2) growth loop as SpaDOT calls it (returns gammas[0], ot_solvers.py:121) and the growth trace
    gamma0 = quiet(ref.compute_transport_map, a, b, dict(cfg), G=None if G is None else G.copy())
    growth = [G0.copy()]
    for _ in range(cfg["growth_iters"]):
        growth.append(quiet(ref.optimal_transport_duality_gap, C=Cn.copy(), G=growth[-1].copy(), **cfg).sum(axis=1))
    last_plan = quiet(ref.optimal_transport_duality_gap, C=Cn.copy(), G=growth[-2].copy(), **cfg)
    # (3) fixed-schedule solver (ot_solvers.py:452)
    plan_v2 = quiet(ref.transport_stablev2, C=Cn.copy(), G=G0.copy(), **cfg)
    out = dict(a=a, b=b, labels_a=la, labels_b=lb, median=med, G=G0,
               gap_row_sums=plan_gap.sum(1), gap_col_sums=plan_gap.sum(0),
               v2_row_sums=plan_v2.sum(1), v2_col_sums=plan_v2.sum(0),
               growth_row_sums=np.stack(growth[1:]),
               table_gap=ot_dense.transition_table(plan_gap, la, lb, 10, 10),
               table_last=ot_dense.transition_table(last_plan, la, lb, 10, 10),
               cfg_keys=np.array(sorted(cfg)), cfg_vals=np.array([float(cfg[k]) for k in sorted(cfg)]))
    assert np.allclose(gamma0, plan_gap, rtol=1e-12, atol=0)  # shipped libot.so is alignment-dependent at 1e-16
    if full_plan:
        out.update(plan_gap=plan_gap, plan_v2=plan_v2)
    else:
        rng = np.random.default_rng(seed + 1)
        ii = rng.integers(0, n, n_samples)
        jj = rng.integers(0, m, n_samples)
        out.update(sample_i=ii, sample_j=jj, plan_gap_samples=plan_gap[ii, jj], plan_v2_samples=plan_v2[ii, jj])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "written;", "plan sum", plan_gap.sum(), "v2-vs-gap", np.abs(plan_v2 - plan_gap).max() / plan_gap.max())


def svgp_case(svgp_mod, name="svgp_small", seed=7):
    g = torch.Generator().manual_seed(seed)
    dt = torch.float64
    n, m, b = 57, 23, 40
    x = torch.rand(n, 2, generator=g, dtype=dt) * 2 - 1
    z = torch.rand(m, 2, generator=g, dtype=dt) * 2 - 1
    out = dict(x=x.numpy(), z=z.numpy())
    for kt in ("Gaussian", "Cauchy", "Quadratic"):
        k = svgp_mod.Kernel(kernel_type=kt, scale=0.1, dtype=dt, device="cpu")
        out["K_" + kt] = k(x, z).numpy()
        out["Kxx_" + kt] = k(x, x).numpy()
    cfg = dict(dtype=dt, device="cpu", kernel_type="Gaussian", kernel_scale=0.1)
    model = svgp_mod.SVGP(cfg, z.numpy(), N_train=n, jitter=1e-2)
    xb = x[:b]
    y = torch.randn(b, generator=g, dtype=dt)
    noise = torch.rand(b, generator=g, dtype=dt) + 0.5
    mean, B, mu_hat, A_hat = model.approximate_posterior_params(xb, xb, y, noise)
    l3, kl = model.variational_loss(xb, y, noise, mu_hat, A_hat)
    out.update(y=y.numpy(), noise=noise.numpy(), post_mean=mean.numpy(), post_var=B.numpy(), mu_hat=mu_hat.numpy(),
               A_hat=A_hat.numpy(), l3=np.array(float(l3)), kl=np.array(float(kl)), N_train=np.array(n), b=np.array(b))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "written; l3", float(l3), "kl", float(kl))


def model_case(name="model_small", seed=5):
    """One forward/backward of the reference SpaDOT model (GATConv stubbed, see reference_loader.load_model)
    on a small synthetic batch with teacher-forced reparameterisation noise."""
    import torch
    from oracle import graph_ref
    ref = reference_loader.load_model()
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    n, genes, b, m = 260, 40, 96, 30
    coords = rng.uniform(-1.5, 1.5, size=(n, 2))
    ei = graph_ref.spatial_edge_index(coords, 6)
    cfg = dict(input_dim=genes, z_dim=8, dtype=torch.float64, device="cpu", svgp_encoder_layers=[32, 16],
               gat_encoder_hidden=12, gat_attention_heads=2, decoder_layers=[16, 32], kernel_type="Gaussian",
               kernel_scale=0.1, timepoints=["t0"])
    dl = dict(inducing_points={"t0": coords[rng.choice(n, m, replace=False)]}, N_train={"t0": n})
    model = ref.SpaDOT(cfg, dl)
    y = torch.from_numpy(rng.normal(size=(n, genes)))
    x = torch.from_numpy(coords)
    noise = [torch.from_numpy(rng.normal(size=(b, 4))) for _ in range(2)]
    it = iter(noise)
    orig = torch.randn_like
    torch.randn_like = lambda t: next(it)
    try:
        recon, skl, gkl, align, final = model.forward(x, y, torch.from_numpy(ei), "t0", b)
    finally:
        torch.randn_like = orig
    loss = 0.1 * recon - 0.7 * skl + 1e-4 * gkl + 0.1 * align
    loss.backward()
    out = dict(coords=coords, y=y.numpy(), edge_index=ei, inducing=dl["inducing_points"]["t0"], b=np.array(b),
               noise0=noise[0].numpy(), noise1=noise[1].numpy(), recon=float(recon), svgp_kl=float(skl), gat_kl=float(gkl),
               align=float(align), final=final.detach().numpy(), loss=float(loss))
    for k, v in model.state_dict().items():
        out["param::" + k] = v.numpy()
    for k, p in model.named_parameters():
        out["grad::" + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "written; losses", float(recon), float(skl), float(gkl), float(align))


TRAIN_CFG = dict(z_dim=8, svgp_encoder_layers=[32, 16], gat_encoder_hidden=12, gat_attention_heads=2, decoder_layers=[16, 32],
                 kernel_type="Gaussian", kernel_scale=0.1, lr=3e-4, lambda1=0.1, beta1=1.0, beta2=1e-4, omiga1=0.1, maxiter=100)


def train_problem(seed=9, sizes=(300, 260), genes=40, m=28, batch=64):
    """Two small timepoints: standardised coordinates, expression, reference-built kNN graph, induced 2-hop batches."""
    from oracle import graph_ref
    rng = np.random.default_rng(seed)
    tps = {}
    for t, n in enumerate(sizes):
        raw = rng.uniform(0, 3000, size=(n, 2))
        loc = (raw - raw.mean(0)) / raw.std(0)                    # StandardScaler per timepoint, _train_utils.py:133-138
        ei = graph_ref.spatial_edge_index(raw, 6)
        batches = [(nodes, lei, ns) for nodes, lei, ns in graph_ref.two_hop_batches_ref(ei, n, batch)]
        tps[f"t{t}"] = dict(loc=loc, y=rng.normal(size=(n, genes)), edge_index=ei, inducing=loc[rng.choice(n, m, replace=False)],
                            batches=batches, n=n)
    return tps


def train_run(model, tps, epochs, noise_seed, device="cpu"):
    """The optimiser loop of train_SpaDOT (utils/_train_utils.py:155-236) before the k-means / OT regularisers switch on:
    AdamW(lr), elbo = lambda1*Recon - beta1(epoch)*SVGP_KL + beta2*GAT_KL + omiga1*alignment, clip_grad_norm_ 0.3, fixed
    timepoint order; reparameterisation noise drawn from one seeded CPU generator (teacher-forced across devices)."""
    from oracle import graph_ref
    gen = torch.Generator().manual_seed(noise_seed)
    noise = lambda t: torch.randn(t.shape, dtype=t.dtype, generator=gen).to(t.device)       # noqa: E731
    opt = torch.optim.AdamW(filter(lambda p: p.requires_grad, model.parameters()), lr=TRAIN_CFG["lr"])
    beta1s = graph_ref.beta_cycle_linear(TRAIN_CFG["maxiter"], stop=TRAIN_CFG["beta1"])
    model.train()
    trace = []
    orig = torch.randn_like
    torch.randn_like = noise
    if hasattr(model, "noise_fn"):
        model.noise_fn = noise
    try:
        for epoch in range(epochs):
            for tp, d in tps.items():
                x_all, y_all = torch.as_tensor(d["loc"], device=device), torch.as_tensor(d["y"], device=device)
                for nodes, lei, ns in d["batches"]:
                    nodes_t, lei_t = torch.as_tensor(nodes, device=device), torch.as_tensor(lei, device=device)
                    recon, skl, gkl, align, _ = model.forward(x_all[nodes_t], y_all[nodes_t], lei_t, tp, ns)
                    elbo = TRAIN_CFG["lambda1"] * recon - beta1s[epoch] * skl + TRAIN_CFG["beta2"] * gkl + TRAIN_CFG["omiga1"] * align
                    opt.zero_grad()
                    elbo.backward()
                    torch.nn.utils.clip_grad_norm_(model.parameters(), 0.3)
                    opt.step()
                    trace.append([float(elbo), float(recon), float(skl), float(gkl), float(align)])
    finally:
        torch.randn_like = orig
    return np.array(trace)


def train_case(name="train_trace_small", seed=9, epochs=3):
    """30 optimiser steps of the reference's OWN model class (GATConv stubbed, see reference_loader.load_model) in the
    reference's training loop: the loss trace north_star asks to reproduce "within 1e-3 relative with the same seed"."""
    ref = reference_loader.load_model()
    tps = train_problem(seed)
    torch.manual_seed(seed)
    cfg = dict(input_dim=40, dtype=torch.float64, device="cpu", timepoints=list(tps), **{k: TRAIN_CFG[k] for k in
               ("z_dim", "svgp_encoder_layers", "gat_encoder_hidden", "gat_attention_heads", "decoder_layers", "kernel_type", "kernel_scale")})
    dl = dict(inducing_points={tp: d["inducing"] for tp, d in tps.items()}, N_train={tp: d["n"] for tp, d in tps.items()})
    model = ref.SpaDOT(cfg, dl)
    out = {"param::" + k: v.clone().numpy() for k, v in model.state_dict().items()}
    trace = train_run(model, tps, epochs, noise_seed=seed + 1)
    out.update(trace=trace, epochs=np.array(epochs), seed=np.array(seed))
    for k, v in model.state_dict().items():
        out["final::" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "written;", len(trace), "steps; elbo", trace[0, 0], "->", trace[-1, 0])


def visium_like_coords(n, seed, jitter=0.35):
    """Hexagonal lattice in pixel units (100 px pitch) with sub-pixel registration jitter, cropped to n spots: the
    geometry of a Visium capture area (ChickenHeart).  The jitter breaks the exact distance ties of an ideal lattice,
    as real full-resolution pixel coordinates do."""
    rng = np.random.default_rng(seed)
    side = int(np.ceil(np.sqrt(n))) + 2
    pts = np.array([[i + 0.5 * (j % 2), j * np.sqrt(3) / 2] for j in range(side) for i in range(side)]) * 100.0
    keep = np.sort(rng.permutation(pts.shape[0])[:n])          # tissue does not cover the whole lattice
    return pts[keep] + rng.normal(0.0, jitter, size=(n, 2))


def graph_case(name, n, seed, kind):
    """edge_index of the reference's own `_Cal_Spatial_Net` + dense_to_sparse (utils/_utils.py:52-100,
    utils/_train_utils.py:69-72) at ChickenHeart sizes; the edge counts printed by the reference must be those of
    examples/ChickenHeart.ipynb:221-230 (4482 @ 747 spots, 23592 @ 1966 spots)."""
    k = int(min(30, 6 * round(n / 1000)))                       # utils/_train_utils.py:69
    coords = visium_like_coords(n, seed) if kind == "visium" else np.random.default_rng(seed).uniform(0, 6000, size=(n, 2))
    ei, log = reference_loader.reference_edge_index(coords, k)
    n_edges = int(log.split("The graph contains ")[1].split(" edges")[0])
    assert n_edges == n * k and ei.shape[1] == n * (k + 1)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), coords=coords, edge_index=ei.astype(np.int32), k=np.array(k),
                        printed_edges=np.array(n_edges))
    print(name, "written;", log.strip().splitlines()[1])


if __name__ == "__main__" and "--train-only" in sys.argv:
    train_case()
    sys.exit(0)


if __name__ == "__main__" and "--graph-only" in sys.argv:
    graph_case("graph_visium_747", 747, 21, "visium")
    graph_case("graph_visium_1966", 1966, 22, "visium")
    graph_case("graph_uniform_1916", 1916, 23, "uniform")
    train_case()
    sys.exit(0)


def sweep_cases(ref):
    """BASELINE configs[4], the corners of the epsilon / unbalance sweep (eps 0.01-0.1, lambda 1-50), at a size the
    reference solves in seconds: pins the oracle where the exponent scale 1/eps is largest and where the near-balanced
    problem needs thousands of iterations."""
    for tag, eps, l1, l2, seed in (("e001_l1_1", 0.01, 1.0, 1.0, 31), ("e001_l1_50", 0.01, 1.0, 50.0, 32), ("e01_l50_50", 0.1, 50.0, 50.0, 33)):
        ot_case(ref, f"ot_sweep_{tag}_120x100_d32", 120, 100, 32, seed, dict(CFG, lambda1=l1, lambda2=l2, epsilon=eps), full_plan=False)


if __name__ == "__main__" and "--sweep-only" in sys.argv:
    assert reference_loader.available(), "reference not mounted"
    sweep_cases(reference_loader.load_ot_solvers())
    sys.exit(0)


if __name__ == "__main__" and "--model-only" in sys.argv:
    model_case()


if __name__ == "__main__" and "--model-only" not in sys.argv:
    assert reference_loader.available(), "reference not mounted"
    ref = reference_loader.load_ot_solvers()
    ot_case(ref, "ot_small_48x61_d6", 48, 61, 6, 11, CFG)
    rng = np.random.default_rng(5)
    ot_case(ref, "ot_growth_90x70_d20", 90, 70, 20, 12, CFG, G=np.exp(rng.normal(0, 1.0, 90)))
    ot_case(ref, "ot_medium_300x411_d20", 300, 411, 20, 1993, CFG, full_plan=False)
    wot_cfg = dict(CFG, lambda1=1.0, lambda2=50.0, epsilon=0.02)
    ot_case(ref, "ot_wotcfg_130x97_d32", 130, 97, 32, 13, wot_cfg, full_plan=False)
    sweep_cases(ref)
    svgp_case(reference_loader.load_svgp())
    model_case()
    graph_case("graph_visium_747", 747, 21, "visium")
    graph_case("graph_visium_1966", 1966, 22, "visium")
    graph_case("graph_uniform_1916", 1916, 23, "uniform")

"""Train inner loop: the model-forward oracle is pinned by a golden batch generated from the reference
model itself (tests/golden/make_golden.py::model_case); the CUDA-backed drop-in must reproduce losses,
latents and every parameter gradient (north_star: loss within 1e-3 relative — here 1e-8)."""
import os

import numpy as np
import pytest
import torch

from oracle import model_ref


def setup(golden_dir, device, cls, **kw):
    z = np.load(os.path.join(golden_dir, "model_small.npz"))
    cfg = dict(input_dim=z["y"].shape[1], z_dim=8, dtype=torch.float64, device=device, svgp_encoder_layers=[32, 16],
               gat_encoder_hidden=12, gat_attention_heads=2, decoder_layers=[16, 32], kernel_type="Gaussian",
               kernel_scale=0.1, timepoints=["t0"])
    dl = dict(inducing_points={"t0": z["inducing"]}, N_train={"t0": z["coords"].shape[0]})
    noise = iter([torch.from_numpy(z["noise0"]).to(device), torch.from_numpy(z["noise1"]).to(device)])
    model = cls(cfg, dl, noise_fn=lambda t: next(noise), **kw).to(device)
    sd = {k[len("param::"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param::")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not [m for m in missing if "svgp" not in m], (missing, unexpected)
    return z, model


def run(z, model, device):
    x = torch.from_numpy(z["coords"]).to(device)
    y = torch.from_numpy(z["y"]).to(device)
    ei = torch.from_numpy(z["edge_index"]).to(device)
    recon, skl, gkl, align, final = model.forward(x, y, ei, "t0", int(z["b"]))
    loss = 0.1 * recon - 0.7 * skl + 1e-4 * gkl + 0.1 * align
    loss.backward()
    return recon, skl, gkl, align, final, loss


def check(z, model, outs, rtol, gtol):
    recon, skl, gkl, align, final, loss = outs
    for got, key in ((recon, "recon"), (skl, "svgp_kl"), (gkl, "gat_kl"), (align, "align"), (loss, "loss")):
        assert float(got) == pytest.approx(float(z[key]), rel=rtol), key
    np.testing.assert_allclose(final.detach().cpu().numpy(), z["final"], rtol=rtol * 10, atol=1e-10)
    for name, p in model.named_parameters():
        want = z["grad::" + name]
        scale = float(np.abs(want).max())        # biases in front of a BatchNorm have gradients that are rounding noise
        assert float(np.abs(p.grad.cpu().numpy() - want).max()) < gtol * scale + 1e-13, name


def test_model_oracle_reproduces_reference_batch(golden_dir):
    z, model = setup(golden_dir, "cpu", model_ref.SpaDOTRef)
    check(z, model, run(z, model, "cpu"), rtol=1e-11, gtol=1e-9)


@pytest.mark.gpu
def test_cuda_model_reproduces_reference_batch(golden_dir):
    from spadot_b200 import model as product
    z, model = setup(golden_dir, torch.device("cuda:0"), product.SpaDOT)
    check(z, model, run(z, model, torch.device("cuda:0")), rtol=1e-9, gtol=1e-6)
    # state_dict keys are the reference's (a saved SpaDOT_model.pth loads unchanged)
    keys = {k[len("param::"):] for k in z.files if k.startswith("param::")}
    assert keys == set(model.state_dict().keys())


@pytest.mark.gpu
def test_all_latent_samples_matches_oracle(golden_dir):
    from spadot_b200 import model as product
    z, model = setup(golden_dir, torch.device("cuda:0"), product.SpaDOT)
    z2, ref = setup(golden_dir, "cpu", model_ref.SpaDOTRef)
    model.eval(), ref.eval()
    got = model.all_latent_samples(z["coords"], z["y"], z["edge_index"], "t0")
    with torch.no_grad():
        x, y = torch.from_numpy(z["coords"]), torch.from_numpy(z["y"])
        q_mu, q_var = ref.SVGPEncoder(y)
        pm = torch.stack([ref.svgp["t0"].approximate_posterior_params(x, x, q_mu[:, l], q_var[:, l])[0] for l in range(4)], dim=1)
        gm, _ = ref.GATEncoder(y, torch.from_numpy(z["edge_index"]))
        want = torch.cat((pm, gm), dim=1).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-8, atol=1e-10)


# ------------------------------------------------------------------ multi-step training parity (north_star: loss curves 1e-3)
def _train_setup(golden_dir, device, cls):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(golden_dir, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    z = np.load(os.path.join(golden_dir, "train_trace_small.npz"))
    tps = mg.train_problem(int(z["seed"]))
    cfg = dict(input_dim=40, dtype=torch.float64, device=device, timepoints=list(tps), **{k: mg.TRAIN_CFG[k] for k in
               ("z_dim", "svgp_encoder_layers", "gat_encoder_hidden", "gat_attention_heads", "decoder_layers", "kernel_type", "kernel_scale")})
    dl = dict(inducing_points={tp: d["inducing"] for tp, d in tps.items()}, N_train={tp: d["n"] for tp, d in tps.items()})
    model = cls(cfg, dl).to(device)
    sd = {k[len("param::"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param::")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not [m for m in missing if "svgp" not in m], (missing, unexpected)
    return mg, z, tps, model


def _check_trace(z, model, trace, rtol):
    want = z["trace"]
    assert trace.shape == want.shape and want.shape[0] >= 28
    rel = np.abs(trace - want) / np.maximum(np.abs(want), 1e-12)
    assert rel.max() < rtol, (rel.max(), np.unravel_index(rel.argmax(), rel.shape))
    for k, v in model.state_dict().items():
        if "final::" + k in z.files and v.dtype.is_floating_point:
            w = z["final::" + k]
            assert float(np.abs(v.cpu().numpy() - w).max()) <= rtol * 10 * max(1.0, float(np.abs(w).max())), k


def test_model_oracle_reproduces_reference_training_trace(golden_dir):
    """oracle/model_ref.py over 30 optimiser steps (3 epochs x 2 timepoints x 5 induced 2-hop batches; AdamW, clip 0.3,
    beta1 cycle, BatchNorm running statistics, the ce > inside_elbo sign rule of SpaDOT.py:77) against the trace the
    reference's own model class produced in the reference's loop (tests/golden/make_golden.py::train_case)."""
    mg, z, tps, model = _train_setup(golden_dir, "cpu", model_ref.SpaDOTRef)
    trace = mg.train_run(model, tps, int(z["epochs"]), noise_seed=int(z["seed"]) + 1)
    _check_trace(z, model, trace, rtol=1e-8)


@pytest.mark.gpu
def test_cuda_model_reproduces_reference_training_trace(golden_dir):
    """The CUDA-backed drop-in model through the same 30 steps.  north_star: "training loss curves within 1e-3 relative
    with the same seed"; the gate here is 1e-6 (measured ~1e-12: fp64 kernels, only summation orders differ)."""
    from spadot_b200 import model as product
    dev = torch.device("cuda:0")
    mg, z, tps, model = _train_setup(golden_dir, dev, product.SpaDOT)
    trace = mg.train_run(model, tps, int(z["epochs"]), noise_seed=int(z["seed"]) + 1, device=dev)
    _check_trace(z, model, trace, rtol=1e-6)

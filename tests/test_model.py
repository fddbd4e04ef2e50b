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

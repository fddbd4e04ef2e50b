"""SVGP prior: the oracle restatement is pinned by the reference-generated golden file (CPU), and the
CUDA-backed drop-in (K1 + cached factorisations + O(b m^2) trace) is checked against both (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import svgp_ref

KTYPES = ["Gaussian", "Cauchy", "Quadratic"]


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "svgp_small.npz"))


def test_oracle_kernel_and_svgp_reproduce_reference(gold):
    x, z = torch.from_numpy(gold["x"]), torch.from_numpy(gold["z"])
    for kt in KTYPES:
        np.testing.assert_allclose(svgp_ref.kernel_forward(x, z, kt, 0.1).numpy(), gold["K_" + kt], rtol=1e-13, atol=1e-15)
    b = int(gold["b"])
    ref = svgp_ref.SVGPRef(gold["z"], int(gold["N_train"]))
    y, noise = torch.from_numpy(gold["y"]), torch.from_numpy(gold["noise"])
    mean, B, mu_hat, A_hat = ref.approximate_posterior_params(x[:b], x[:b], y, noise)
    np.testing.assert_allclose(mean.numpy(), gold["post_mean"], rtol=1e-10)
    np.testing.assert_allclose(B.numpy(), gold["post_var"], rtol=1e-10)
    np.testing.assert_allclose(A_hat.numpy(), gold["A_hat"], rtol=1e-9, atol=1e-12)
    l3, kl = ref.variational_loss(x[:b], y, noise, mu_hat, A_hat)
    assert float(l3) == pytest.approx(float(gold["l3"]), rel=1e-10)
    assert float(kl) == pytest.approx(float(gold["kl"]), rel=1e-9)


def test_unknown_kernel_type_raises_like_reference():
    with pytest.raises(UnboundLocalError):
        svgp_ref.kernel_forward(torch.zeros(2, 2, dtype=torch.float64), torch.zeros(2, 2, dtype=torch.float64), "Matern")
    from spadot_b200 import svgp
    with pytest.raises(UnboundLocalError):
        svgp.kernel_block(torch.zeros(2, 2), torch.zeros(2, 2), "Matern", 0.1)


@pytest.mark.gpu
def test_kernel_blocks_match_golden_on_gpu(gold):
    from spadot_b200 import svgp
    dev = torch.device("cuda:0")
    x, z = torch.from_numpy(gold["x"]).to(dev), torch.from_numpy(gold["z"]).to(dev)
    for kt in KTYPES:
        k = svgp.Kernel(kernel_type=kt, scale=0.1, dtype=torch.float64, device=dev)
        np.testing.assert_allclose(k(x, z).cpu().numpy(), gold["K_" + kt], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(k(x, x).cpu().numpy(), gold["Kxx_" + kt], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(svgp.kernel_diag(x, x, kt, 0.1).cpu().numpy(), np.ones(x.shape[0]), rtol=0, atol=1e-15)
        kj = svgp.kernel_block(z, z, kt, 0.1, jitter=0.25).cpu().numpy()
        np.testing.assert_allclose(kj, gold["K_" + kt][:0].shape and svgp_ref.kernel_forward(z.cpu(), z.cpu(), kt, 0.1).numpy() + 0.25 * np.eye(z.shape[0]), rtol=1e-12, atol=1e-13)


@pytest.mark.gpu
def test_svgp_dropin_matches_golden_and_oracle_with_grads(gold):
    from spadot_b200 import svgp
    dev = torch.device("cuda:0")
    cfg = dict(dtype=torch.float64, device=dev, kernel_type="Gaussian", kernel_scale=0.1)
    b, N = int(gold["b"]), int(gold["N_train"])
    model = svgp.SVGP(cfg, gold["z"], N_train=N, jitter=1e-2)
    x = torch.from_numpy(gold["x"]).to(dev)[:b]
    y = torch.from_numpy(gold["y"]).to(dev).requires_grad_(True)
    noise = torch.from_numpy(gold["noise"]).to(dev).requires_grad_(True)
    mean, B, mu_hat, A_hat = model.approximate_posterior_params(x, x, y, noise)
    np.testing.assert_allclose(mean.detach().cpu().numpy(), gold["post_mean"], rtol=1e-9)
    np.testing.assert_allclose(B.detach().cpu().numpy(), gold["post_var"], rtol=1e-9)
    np.testing.assert_allclose(mu_hat.detach().cpu().numpy(), gold["mu_hat"], rtol=1e-9, atol=1e-12)
    l3, kl = model.variational_loss(x, y, noise, mu_hat, A_hat)
    assert float(l3) == pytest.approx(float(gold["l3"]), rel=1e-9)
    assert float(kl) == pytest.approx(float(gold["kl"]), rel=1e-8)
    # gradients through y / noise must equal the oracle's autograd
    (l3 + kl).backward()
    ref = svgp_ref.SVGPRef(gold["z"], N)
    yc = torch.from_numpy(gold["y"]).requires_grad_(True)
    nc = torch.from_numpy(gold["noise"]).requires_grad_(True)
    xc = torch.from_numpy(gold["x"])[:b]
    _, _, mu_c, A_c = ref.approximate_posterior_params(xc, xc, yc, nc)
    l3c, klc = ref.variational_loss(xc, yc, nc, mu_c, A_c)
    (l3c + klc).backward()
    np.testing.assert_allclose(y.grad.cpu().numpy(), yc.grad.numpy(), rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(noise.grad.cpu().numpy(), nc.grad.numpy(), rtol=1e-7, atol=1e-10)


@pytest.mark.gpu
def test_svgp_chickenheart_shape_and_diag_only_scale():
    """b=512, m=389 (ChickenHeart tp 3): O(b m^2) trace equals the reference's (b,m,m) product; a
    100k-point diagonal costs 100k outputs, not an 80 GB block."""
    from spadot_b200 import svgp
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    z = torch.rand(389, 2, generator=g, dtype=torch.float64) * 4 - 2
    x = torch.rand(512, 2, generator=g, dtype=torch.float64) * 4 - 2
    y = torch.randn(512, generator=g, dtype=torch.float64)
    noise = torch.rand(512, generator=g, dtype=torch.float64) + 0.5
    cfg = dict(dtype=torch.float64, device=dev, kernel_type="Gaussian", kernel_scale=0.1)
    model = svgp.SVGP(cfg, z.numpy(), N_train=1916)
    mean, B, mu_hat, A_hat = model.approximate_posterior_params(x.to(dev), x.to(dev), y.to(dev), noise.to(dev))
    l3, kl = model.variational_loss(x.to(dev), y.to(dev), noise.to(dev), mu_hat, A_hat)
    ref = svgp_ref.SVGPRef(z.numpy(), 1916)
    mean_c, B_c, mu_c, A_c = ref.approximate_posterior_params(x, x, y, noise)
    l3c, klc = ref.variational_loss(x, y, noise, mu_c, A_c)
    np.testing.assert_allclose(B.cpu().numpy(), B_c.numpy(), rtol=1e-7, atol=1e-10)
    assert float(l3) == pytest.approx(float(l3c), rel=1e-8)
    assert float(kl) == pytest.approx(float(klc), rel=1e-7)
    big = torch.rand(100_000, 2, dtype=torch.float64, device=dev)
    d = model.kernel_matrix(big, big, diag_only=True)
    assert d.shape == (100_000,) and float((d - 1).abs().max()) == 0.0


@pytest.mark.gpu
def test_batched_latent_dims_equal_the_per_dimension_loop():
    from spadot_b200 import svgp
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    z = torch.rand(135, 2, generator=g, dtype=torch.float64) * 4 - 2
    x = (torch.rand(300, 2, generator=g, dtype=torch.float64) * 4 - 2).to(dev)
    Y = torch.randn(300, 6, generator=g, dtype=torch.float64).to(dev).requires_grad_(True)
    NZ = (torch.rand(300, 6, generator=g, dtype=torch.float64) + 0.5).to(dev).requires_grad_(True)
    cfg = dict(dtype=torch.float64, device=dev, kernel_type="Cauchy", kernel_scale=0.3)
    model = svgp.SVGP(cfg, z.numpy(), N_train=747)
    mean, var, L3, KL = model.posterior_and_loss_all_dims(x, Y, NZ)
    (L3.sum() + KL.sum() + (mean * var).sum()).backward()
    gY, gN = Y.grad.clone(), NZ.grad.clone()
    Y.grad = NZ.grad = None
    tot = 0
    for l in range(6):
        mean_l, var_l, mu_hat, A_hat = model.approximate_posterior_params(x, x, Y[:, l], NZ[:, l])
        l3, kl = model.variational_loss(x, Y[:, l], NZ[:, l], mu_hat, A_hat)
        np.testing.assert_allclose(mean[:, l].detach().cpu().numpy(), mean_l.detach().cpu().numpy(), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(var[:, l].detach().cpu().numpy(), var_l.detach().cpu().numpy(), rtol=1e-9, atol=1e-12)
        assert float(L3[l]) == pytest.approx(float(l3), rel=1e-10)
        assert float(KL[l]) == pytest.approx(float(kl), rel=1e-9)
        tot = tot + l3 + kl + (mean_l * var_l).sum()
    tot.backward()
    np.testing.assert_allclose(gY.cpu().numpy(), Y.grad.cpu().numpy(), rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(gN.cpu().numpy(), NZ.grad.cpu().numpy(), rtol=1e-7, atol=1e-10)

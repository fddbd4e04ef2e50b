"""Drop-in SVGP prior (SpaDOT/model/svgp.py) with fused CUDA kernel-block construction (K1).

Same public surface as the reference: `Kernel(kernel_type, scale, dtype, device).forward(x, y)`,
`SVGP(model_config, inducing_points, N_train, jitter)` with `kernel_matrix`,
`approximate_posterior_params`, `variational_loss`.  What changes underneath:

* kernel blocks come from `sdb_kernel_block_f64` / `sdb_kernel_diag_f64` (no cdist, no n x n block for
  a diagonal, jitter fused) — they carry no gradient in the reference either (svgp.py:24-25,114);
* `K_mm`, `(K_mm + jI)^-1` and `log det(K_mm + jI)` are computed once per SVGP instead of
  2 x z/2 times per mini-batch (SpaDOT.py:57-62; svgp.py:49-50,64-65,87);
* the (b,m,m) batched product of `_compute_l3_term` (svgp.py:96-104) is the O(b m^2) identity
  trace_i = w_i^T A_hat w_i, w_i = K_mm^-1 k_i;
* posterior variances use row-wise quadratic forms instead of n x n products (svgp.py:78-79).

m x m inverses / Choleskys stay on cuSOLVER in fp64 (condition numbers 1e3-1e5).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib

# the reference evaluates b * torch.log(2 * torch.tensor(torch.pi)) in float32 (svgp.py:103)
LOG_2PI_F32 = torch.log(2 * torch.tensor(torch.pi, dtype=torch.float32))

_KERNEL_IDS = {"Gaussian": 0, "Cauchy": 1, "Quadratic": 2}


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def kernel_block(x, y, kernel_type, scale, jitter=0.0):
    """k(|x_i - y_j|^2) (+ jitter on the diagonal) as an (n, m) float64 CUDA tensor; no autograd."""
    if kernel_type not in _KERNEL_IDS:
        # the reference falls through its three `if`s and fails on `return res` (svgp.py:118-125)
        raise UnboundLocalError("cannot access local variable 'res' where it is not associated with a value")
    _lib.require_device()
    x = x.detach().to(torch.float64).contiguous()
    y = y.detach().to(torch.float64).contiguous()
    if not x.is_cuda or not y.is_cuda:
        raise RuntimeError("spadot_b200.svgp needs CUDA tensors; there is no CPU fallback")
    n, dim = x.shape
    m = y.shape[0]
    out = torch.empty((n, m), dtype=torch.float64, device=x.device)
    _lib.call("sdb_kernel_block_f64", x.data_ptr(), n, y.data_ptr(), m, dim, _KERNEL_IDS[kernel_type], float(scale),
              float(jitter), out.data_ptr(), m, _stream(x))
    return out


def kernel_diag(x, y, kernel_type, scale):
    if kernel_type not in _KERNEL_IDS:
        raise UnboundLocalError("cannot access local variable 'res' where it is not associated with a value")
    _lib.require_device()
    x = x.detach().to(torch.float64).contiguous()
    y = y.detach().to(torch.float64).contiguous()
    out = torch.empty(x.shape[0], dtype=torch.float64, device=x.device)
    _lib.call("sdb_kernel_diag_f64", x.data_ptr(), y.data_ptr(), x.shape[0], x.shape[1], _KERNEL_IDS[kernel_type],
              float(scale), out.data_ptr(), _stream(x))
    return out


class _QuadFormRows(torch.autograd.Function):
    """q_i = w_i^T A w_i with gradient w.r.t. A only (W = K_nm K_mm^-1 is constant data)."""

    @staticmethod
    def forward(ctx, W, A):
        W = W.contiguous()
        Ad = A.detach().contiguous()
        out = torch.empty(W.shape[0], dtype=torch.float64, device=W.device)
        _lib.call("sdb_quad_form_rows_f64", W.data_ptr(), Ad.data_ptr(), W.shape[0], W.shape[1], out.data_ptr(), _stream(W))
        ctx.save_for_backward(W)
        return out

    @staticmethod
    def backward(ctx, grad):
        (W,) = ctx.saved_tensors
        return None, torch.matmul(W.T * grad[None, :], W)          # dq_i/dA = w_i w_i^T


class Kernel(nn.Module):
    def __init__(self, kernel_type="Gaussian", scale=0.1, dtype=torch.float64, device="cpu"):
        super().__init__()
        self.kernel_type = kernel_type
        self.scale = torch.tensor([scale], dtype=dtype, device=device)
        self._scale_value = float(scale)

    def forward(self, x, y):
        return kernel_block(x, y, self.kernel_type, self._scale_value).to(self.scale.dtype)


class SVGP(nn.Module):
    def __init__(self, model_config, inducing_points, N_train, jitter=1e-2):
        super().__init__()
        self.N_train = N_train
        self.jitter = jitter
        self.inducing_index_points = torch.tensor(inducing_points, dtype=model_config["dtype"]).to(model_config["device"])
        self.kernel = Kernel(kernel_type=model_config["kernel_type"], scale=model_config["kernel_scale"],
                             dtype=model_config["dtype"], device=model_config["device"])
        # the reference adds jitter * torch.eye(...) with a float32 identity (svgp.py:37-41)
        self._jit = float(torch.tensor(jitter, dtype=torch.get_default_dtype()))
        self._cache = None

    # ------------------------------------------------------------------ reference surface
    def _add_diagonal_jitter(self, matrix, jitter):
        idx = torch.arange(matrix.size(-1), device=matrix.device)
        out = matrix.clone()
        out[..., idx, idx] += float(torch.tensor(jitter, dtype=torch.get_default_dtype()))
        return out

    def kernel_matrix(self, x, y, diag_only=False):
        if diag_only:
            return kernel_diag(x, y, self.kernel.kernel_type, self.kernel._scale_value)
        return self.kernel(x, y)

    # ------------------------------------------------------------------ cached m x m pieces
    def _mm(self):
        if self._cache is None:
            z = self.inducing_index_points
            kt, sc = self.kernel.kernel_type, self.kernel._scale_value
            K_mm = kernel_block(z, z, kt, sc)
            K_mm_j = kernel_block(z, z, kt, sc, jitter=self._jit)            # jitter fused into the block
            K_mm_inv = torch.linalg.inv(K_mm_j)
            chol = torch.linalg.cholesky(K_mm_j)
            log_det = 2 * torch.sum(torch.log(torch.diagonal(chol)))
            self._cache = (K_mm, K_mm_inv, log_det)
        return self._cache

    def approximate_posterior_params(self, index_points_test, index_points_train, y, noise):
        """svgp.py:62-84."""
        b = index_points_train.shape[0]
        K_mm, K_mm_inv, _ = self._mm()
        z = self.inducing_index_points
        K_xx = self.kernel_matrix(index_points_test, index_points_test, diag_only=True)
        K_xm = self.kernel_matrix(index_points_test, z)
        same = index_points_test is index_points_train
        K_nm = K_xm if same else self.kernel_matrix(index_points_train, z)
        K_mn = K_nm.T
        sigma_l = K_mm + (self.N_train / b) * torch.matmul(K_mn, K_nm / noise[:, None])
        sigma_l_inv = torch.linalg.inv(self._add_diagonal_jitter(sigma_l, self.jitter))
        rhs = torch.matmul(K_mn, y / noise)
        mean_vector = (self.N_train / b) * torch.matmul(K_xm, torch.matmul(sigma_l_inv, rhs))
        # diag(-K_xm K_mm^-1 K_mx + K_xm Sigma_l^-1 K_mx) as row-wise quadratic forms (no n x n product)
        B = K_xx + torch.sum(torch.matmul(K_xm, sigma_l_inv - K_mm_inv) * K_xm, dim=1)
        mu_hat = (self.N_train / b) * torch.matmul(torch.matmul(K_mm, torch.matmul(sigma_l_inv, K_mn)), y / noise)
        A_hat = torch.matmul(K_mm, torch.matmul(sigma_l_inv, K_mm))
        return mean_vector, B, mu_hat, A_hat

    def variational_loss(self, x, y, noise, mu_hat, A_hat):
        """svgp.py:47-60 (+ :86-104)."""
        b, m = x.shape[0], self.inducing_index_points.shape[0]
        K_mm, K_mm_inv, K_mm_log_det = self._mm()
        K_nn = self.kernel_matrix(x, x, diag_only=True)
        K_nm = self.kernel_matrix(x, self.inducing_index_points)
        W = torch.matmul(K_nm, K_mm_inv)                                       # rows w_i = K_mm^-1 k_i
        mean_vector = torch.matmul(W, mu_hat)
        S_chol = torch.linalg.cholesky(self._add_diagonal_jitter(A_hat, self.jitter))
        S_log_det = 2 * torch.sum(torch.log(torch.diagonal(S_chol)))
        KL_term = 0.5 * (K_mm_log_det - S_log_det - m + torch.sum(K_mm_inv * A_hat.T)
                         + torch.sum(mu_hat * torch.matmul(K_mm_inv, mu_hat)))
        precision = 1 / noise
        K_tilde_terms = precision * (K_nn - torch.sum(W * K_nm, dim=1))
        trace_terms = precision * _QuadFormRows.apply(W, A_hat)
        L_3_sum_term = -0.5 * (torch.sum(K_tilde_terms) + torch.sum(trace_terms) + torch.sum(torch.log(noise))
                               + float(b * LOG_2PI_F32) + torch.sum(precision * (y - mean_vector) ** 2))
        return L_3_sum_term, KL_term

    # ------------------------------------------------------------------ all latent dimensions at once
    def posterior_and_loss_all_dims(self, x, Y, noise):
        """The z/2-iteration loop of SpaDOT.forward (SpaDOT/model/SpaDOT.py:57-66) as one batched evaluation.

        x (b,2) coordinates of the batch, Y (b,L) and noise (b,L) the encoder means / variances of the L
        GP latent dimensions.  Returns (mean (b,L), var (b,L), L3 (L,), KL (L,)) equal to calling
        approximate_posterior_params(x, x, Y[:,l], noise[:,l]) + variational_loss(...) for each l: the kernel
        blocks are shared by every l, and the L Sigma_l inverses / Choleskys run as one batched cuSOLVER call."""
        b, L = Y.shape
        m = self.inducing_index_points.shape[0]
        K_mm, K_mm_inv, K_mm_log_det = self._mm()
        K_nn = self.kernel_matrix(x, x, diag_only=True)
        K_nm = self.kernel_matrix(x, self.inducing_index_points)
        K_mn = K_nm.T
        scale = self.N_train / b
        prec = 1 / noise                                                       # (b,L)
        sigma_l = K_mm[None] + scale * torch.matmul(K_mn[None] * prec.T[:, None, :], K_nm)     # (L,m,m)
        sigma_l_inv = torch.linalg.inv(self._add_diagonal_jitter(sigma_l, self.jitter))
        rhs = torch.matmul(K_mn, Y * prec)                                     # (m,L)
        sol = torch.matmul(sigma_l_inv, rhs.T[:, :, None])[:, :, 0]            # (L,m)
        mean = scale * torch.matmul(K_nm, sol.T)                               # (b,L)
        var = K_nn[:, None] + torch.sum(torch.matmul(K_nm[None], sigma_l_inv - K_mm_inv[None]) * K_nm[None], dim=2).T
        mu_hat = scale * torch.matmul(K_mm, sol.T).T                           # (L,m)
        A_hat = torch.matmul(K_mm[None], torch.matmul(sigma_l_inv, K_mm[None]))  # (L,m,m)
        W = torch.matmul(K_nm, K_mm_inv)                                       # (b,m)
        mean_vector = torch.matmul(W, mu_hat.T)                                # (b,L)
        S_chol = torch.linalg.cholesky(self._add_diagonal_jitter(A_hat, self.jitter))
        S_log_det = 2 * torch.sum(torch.log(torch.diagonal(S_chol, dim1=1, dim2=2)), dim=1)
        KL = 0.5 * (K_mm_log_det - S_log_det - m + torch.sum(K_mm_inv[None] * A_hat.transpose(1, 2), dim=(1, 2))
                    + torch.sum(mu_hat * torch.matmul(mu_hat, K_mm_inv.T), dim=1))
        K_tilde = prec * (K_nn - torch.sum(W * K_nm, dim=1))[:, None]
        trace_terms = prec * torch.sum(torch.matmul(W[None], A_hat) * W[None], dim=2).T
        L3 = -0.5 * (torch.sum(K_tilde, dim=0) + torch.sum(trace_terms, dim=0) + torch.sum(torch.log(noise), dim=0)
                     + float(b * LOG_2PI_F32) + torch.sum(prec * (Y - mean_vector) ** 2, dim=0))
        return mean, var, L3, KL

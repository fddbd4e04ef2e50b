"""Drop-in `SpaDOT` model (SpaDOT/model/SpaDOT.py) wired to the CUDA-backed SVGP prior and GAT encoder.

Same constructor, `forward(x, y, edge_index, tp, batch_size)` return tuple, `all_latent_samples` and
parameter names as the reference, so `SpaDOT_model.pth` state_dicts load unchanged and
`utils/_train_utils.train_SpaDOT` can drive it as is.  Differences are all inside the hot path:

* the z/2-iteration SVGP loop (SpaDOT.py:57-66) is one batched evaluation
  (`svgp.SVGP.posterior_and_loss_all_dims`: shared kernel blocks from K1, cached K_mm factorisations,
  O(b m^2) trace, one batched cuSOLVER call for the Sigma_l);
* the GAT encoder runs the fused edge-softmax kernels (K2/K2b);
* the `ce_term.item() > inside_elbo.item()` host sync (SpaDOT.py:77) is the sync-free `-|diff|`, which is
  the same value and the same gradient.

SVGPEncoder / Decoder are the reference's plain MLPs (cuBLAS); they are callers of the path, not the path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .gat import GATEncoder
from .svgp import SVGP


class SVGPEncoder(nn.Module):                         # SpaDOT/model/encoder.py:7-34
    def __init__(self, input_dim, SVGP_z_dim, hidden_dims):
        super().__init__()
        layers = [input_dim] + hidden_dims
        net = []
        for i in range(1, len(layers)):
            lin = nn.Linear(layers[i - 1], layers[i])
            nn.init.xavier_uniform_(lin.weight)
            net += [lin, nn.BatchNorm1d(layers[i]), nn.LeakyReLU()]
        self.SVGP_encoder_net = nn.Sequential(*net)
        self.SVGP_fc = nn.Linear(hidden_dims[-1], SVGP_z_dim * 2)
        nn.init.xavier_uniform_(self.SVGP_fc.weight)

    def forward(self, x):
        mu, logvar = torch.chunk(self.SVGP_fc(self.SVGP_encoder_net(x)), 2, dim=1)
        return mu, torch.exp(logvar)


class Decoder(nn.Module):                             # SpaDOT/model/decoder.py:3-20
    def __init__(self, input_dim, z_dim, decoder_layers):
        super().__init__()
        layers = [z_dim] + decoder_layers + [input_dim]
        net = []
        for i in range(1, len(layers) - 1):
            lin = nn.Linear(layers[i - 1], layers[i])
            nn.init.xavier_uniform_(lin.weight)
            net += [lin, nn.LayerNorm(layers[i]), nn.LeakyReLU()]
        net.append(nn.Linear(layers[-2], layers[-1]))
        self.decoder_net = nn.Sequential(*net)

    def forward(self, latent_sample):
        return self.decoder_net(latent_sample)


class SpaDOT(nn.Module):
    def __init__(self, model_config, dataloader_dict, noise_fn=None):
        super().__init__()
        self.input_dim = model_config["input_dim"]
        self.SVGP_z_dim = model_config["z_dim"] // 2
        self.GAT_z_dim = model_config["z_dim"] // 2
        self.dtype = model_config["dtype"]
        self.device = model_config["device"]
        self.SVGPEncoder = SVGPEncoder(self.input_dim, self.SVGP_z_dim, model_config["svgp_encoder_layers"]).to(dtype=self.dtype)
        self.GATEncoder = GATEncoder(self.input_dim, self.GAT_z_dim, model_config["gat_encoder_hidden"],
                                     model_config["gat_attention_heads"]).to(dtype=self.dtype)
        self.decoder = Decoder(self.input_dim, self.SVGP_z_dim + self.GAT_z_dim, model_config["decoder_layers"]).to(dtype=self.dtype)
        self.svgp_dict = nn.ModuleDict({
            str(tp): SVGP(model_config=model_config, inducing_points=dataloader_dict["inducing_points"][tp],
                          N_train=dataloader_dict["N_train"][tp]).to(model_config["device"])
            for tp in model_config["timepoints"]})
        self.gammas = {}
        self.kmeans_center_dict = {}
        self.kmeans_cluster_dict = {}
        self.kmeans_index_dict = {}
        self.noise_fn = noise_fn or torch.randn_like

    def forward(self, x, y, edge_index, tp, batch_size):
        svgp = self.svgp_dict[str(tp)]
        q_mu, q_var = self.SVGPEncoder(y[:batch_size])
        p_m, p_v, rec, kl = svgp.posterior_and_loss_all_dims(x[:batch_size], q_mu, q_var)
        inside_elbo = torch.sum(rec) - (batch_size / svgp.N_train) * torch.sum(kl)
        ce_term = torch.sum(self._gauss_cross_entropy(p_m, p_v, q_mu, q_var))
        diff = ce_term - inside_elbo
        SVGP_KL = -torch.abs(diff) / self.SVGP_z_dim          # SpaDOT.py:77 without the .item() round trip
        SVGP_latent_sample = p_m + self.noise_fn(p_m) * torch.sqrt(p_v)

        # only the seeds' rows are used (SpaDOT.py:83-84): the encoder computes each layer on the rows the next one reads
        GAT_m, GAT_v = self.GATEncoder(y, edge_index, pos=x, n_out=batch_size)
        GAT_m, GAT_v = GAT_m[:batch_size, :], GAT_v[:batch_size, :]
        GAT_latent_sample = GAT_m + self.noise_fn(GAT_m) * torch.sqrt(GAT_v)
        GAT_KL = -0.5 * torch.sum(1 + torch.log(GAT_v) - GAT_m.pow(2) - GAT_v) / self.GAT_z_dim

        final_latent = torch.cat([SVGP_latent_sample, GAT_latent_sample], dim=1)
        recon_loss = torch.sum((y[:batch_size] - self.decoder(final_latent)) ** 2) / self.input_dim
        alignment_loss = F.mse_loss(SVGP_latent_sample.norm(dim=1) / self.SVGP_z_dim,
                                    GAT_latent_sample.norm(dim=1) / self.GAT_z_dim, reduction="sum")
        return recon_loss, SVGP_KL, GAT_KL, alignment_loss, final_latent

    def all_latent_samples(self, X, Y, edge_index, tp):
        """SpaDOT.py:96-123: posterior means of every spot of one timepoint (no n x n temporaries)."""
        X = torch.as_tensor(X, dtype=self.dtype, device=self.device)
        Y = torch.as_tensor(Y, dtype=self.dtype, device=self.device)
        edge_index = torch.as_tensor(edge_index, dtype=torch.long, device=self.device)
        q_mu, q_var = self.SVGPEncoder(Y)
        p_m, _, _, _ = self.svgp_dict[str(tp)].posterior_and_loss_all_dims(X, q_mu, q_var)
        GAT_m, _ = self.GATEncoder(Y, edge_index, pos=X)
        return torch.cat((p_m, GAT_m), dim=1).data.cpu().detach().numpy()

    def _gauss_cross_entropy(self, mu1, var1, mu2, var2):
        term0 = 1.8378770664093453
        return -0.5 * (term0 + torch.log(var2) + (var1 + mu1 ** 2 - 2 * mu1 * mu2 + mu2 ** 2) / var2)

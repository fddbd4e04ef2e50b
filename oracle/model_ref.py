"""torch restatement of the reference's model forward pass.  TEST INFRASTRUCTURE ONLY.

Follows SpaDOT/model/SpaDOT.py:9-142, SpaDOT/model/encoder.py:7-34 (SVGPEncoder) and
SpaDOT/model/decoder.py:3-20 line by line, on top of oracle/svgp_ref.py (pinned) and
oracle/gat_ref.py (PyG semantics, parity unpinned).  Pinned by tests/golden/model_small.npz, generated
by importing the reference model itself with only torch_geometric.GATConv stubbed
(oracle/reference_loader.load_model).  `noise_fn` replaces torch.randn_like so that tests can
teacher-force the reparameterisation noise across devices.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import gat_ref, svgp_ref


class SVGPEncoderRef(nn.Module):                      # encoder.py:7-34
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


class DecoderRef(nn.Module):                          # decoder.py:3-20
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

    def forward(self, z):
        return self.decoder_net(z)


def gauss_cross_entropy(mu1, var1, mu2, var2):        # SpaDOT.py:125-142
    return -0.5 * (1.8378770664093453 + torch.log(var2) + (var1 + mu1 ** 2 - 2 * mu1 * mu2 + mu2 ** 2) / var2)


class SpaDOTRef(nn.Module):
    def __init__(self, model_config, dataloader_dict, noise_fn=None):
        super().__init__()
        self.input_dim = model_config["input_dim"]
        self.SVGP_z_dim = self.GAT_z_dim = model_config["z_dim"] // 2
        dt = model_config["dtype"]
        self.SVGPEncoder = SVGPEncoderRef(self.input_dim, self.SVGP_z_dim, model_config["svgp_encoder_layers"]).to(dt)
        self.GATEncoder = gat_ref.GATEncoderRef(self.input_dim, self.GAT_z_dim, model_config["gat_encoder_hidden"],
                                                model_config["gat_attention_heads"]).to(dt)
        self.decoder = DecoderRef(self.input_dim, self.SVGP_z_dim + self.GAT_z_dim, model_config["decoder_layers"]).to(dt)
        self.svgp = {str(tp): svgp_ref.SVGPRef(dataloader_dict["inducing_points"][tp], dataloader_dict["N_train"][tp],
                                               model_config["kernel_type"], model_config["kernel_scale"]).to(model_config["device"])
                     for tp in model_config["timepoints"]}
        self.noise_fn = noise_fn or torch.randn_like

    def forward(self, x, y, edge_index, tp, batch_size):      # SpaDOT.py:52-94
        sv = self.svgp[str(tp)]
        q_mu, q_var = self.SVGPEncoder(y[:batch_size])
        rec, kl, pm, pv = [], [], [], []
        for l in range(self.SVGP_z_dim):
            m_l, v_l, mu_hat, A_hat = sv.approximate_posterior_params(x[:batch_size], x[:batch_size], q_mu[:, l], q_var[:, l])
            r_l, k_l = sv.variational_loss(x[:batch_size], q_mu[:, l], q_var[:, l], mu_hat, A_hat)
            rec.append(r_l), kl.append(k_l), pm.append(m_l), pv.append(v_l)
        inside_elbo = torch.sum(torch.stack(rec)) - (batch_size / sv.N_train) * torch.sum(torch.stack(kl))
        p_m, p_v = torch.stack(pm, dim=1), torch.stack(pv, dim=1)
        ce = torch.sum(gauss_cross_entropy(p_m, p_v, q_mu, q_var))
        diff = ce - inside_elbo
        SVGP_KL = (-diff if ce.item() > inside_elbo.item() else diff) / self.SVGP_z_dim
        svgp_sample = p_m + self.noise_fn(p_m) * torch.sqrt(p_v)
        g_m, g_v = self.GATEncoder(y, edge_index)
        g_m, g_v = g_m[:batch_size], g_v[:batch_size]
        gat_sample = g_m + self.noise_fn(g_m) * torch.sqrt(g_v)
        GAT_KL = -0.5 * torch.sum(1 + torch.log(g_v) - g_m.pow(2) - g_v) / self.GAT_z_dim
        final = torch.cat([svgp_sample, gat_sample], dim=1)
        recon = torch.sum((y[:batch_size] - self.decoder(final)) ** 2) / self.input_dim
        align = F.mse_loss(svgp_sample.norm(dim=1) / self.SVGP_z_dim, gat_sample.norm(dim=1) / self.GAT_z_dim, reduction="sum")
        return recon, SVGP_KL, GAT_KL, align, final

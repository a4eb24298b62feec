"""DMM proposals and decoder: same classes, constructor arguments, parameter names (state_dict keys) and
forward() contract as apgs/dmm/models.py (Enc_rws_mu :10-55, Enc_apg_local :57-107, Enc_apg_mu :109-151,
Decoder :155-210).  The small MLPs and the random draws stay in PyTorch; NSS pooling and every log-density /
likelihood / weight are in csrc/dmm.cu.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.distributions.beta import Beta
from torch.distributions.normal import Normal
from torch.distributions.one_hot_categorical import OneHotCategorical as cat

from ..rng import normal_sample
from ..trace import Trace
from . import kernels


def _x_inst(ob):
    return ob[0] if ob.dim() == 4 else ob


def needs_grad(*modules):
    """True when any parameter of the given modules takes part in autograd (training mode)."""
    return torch.is_grad_enabled() and any(p.requires_grad for m in modules for p in m.parameters())


def _split_linear(lin, x, z):
    """Linear(cat(x, z)) = Linear_x(x) (per instance, broadcast over S) + Linear_z(z), without the S-replicated cat."""
    D = x.shape[-1]
    return F.linear(z, lin.weight[:, D:]) + F.linear(x, lin.weight[:, :D], lin.bias)


class _EncMu(nn.Module):
    def _build(self, d_in, K, D, num_hidden, num_nss):
        self.nss1 = nn.Sequential(nn.Linear(d_in, num_hidden), nn.Tanh(), nn.Linear(num_hidden, num_nss))
        self.nss2 = nn.Sequential(nn.Linear(d_in, num_hidden), nn.Tanh(), nn.Linear(num_hidden, K), nn.Softmax(-1))
        self.mean_mu = nn.Sequential(nn.Linear(num_nss + 2 * D, num_hidden), nn.Tanh(), nn.Linear(num_hidden, D))
        self.mean_log_sigma = nn.Sequential(nn.Linear(num_nss + 2 * D, num_hidden), nn.Tanh(), nn.Linear(num_hidden, D))

    fused_heads = True    # evaluation path: mean_mu / mean_log_sigma (+ exp) in one launch (csrc/dmm_heads.cu)

    def _mu_heads(self, nss_prior):
        """(q_mu, q_sigma) from cat(nss, priors) (models.py:43-44)"""
        if (self.fused_heads and nss_prior.is_cuda and nss_prior.dtype == torch.float32 and not needs_grad(self)
                and not (torch.is_grad_enabled() and nss_prior.requires_grad)
                and kernels.mu_heads_supported(self.mean_mu, self.mean_log_sigma)):
            return kernels.mu_heads(nss_prior, self.mean_mu, self.mean_log_sigma)
        return self.mean_mu(nss_prior), self.mean_log_sigma(nss_prior).exp()

    def _heads(self, nss1, nss2, priors):
        nss_prior = kernels.nss_pool(nss1, nss2, priors[0], priors[1])          # D1
        return self._mu_heads(nss_prior)

    def _heads_torch(self, nss1, nss2, priors, EPS=1e-8):
        """D1 + heads with PyTorch ops only (autograd graph for training; models.py:134-139)."""
        num = torch.einsum('sbnj,sbnk->sbkj', nss1, nss2)
        nss = num / (nss2.sum(2).unsqueeze(-1) + EPS)
        S, B, K, _ = nss.shape
        nss_prior = torch.cat((nss, priors[0].expand(S, B, K, -1), priors[1].expand(S, B, K, -1)), -1)
        return self.mean_mu(nss_prior), self.mean_log_sigma(nss_prior).exp()

    @staticmethod
    def _trace(q_mu_mu, q_mu_sigma, sampled, mu_old):
        q = Trace()
        if sampled:
            mu = normal_sample(q_mu_mu, q_mu_sigma)
            q.normal(q_mu_mu, q_mu_sigma, value=mu, name='means')
        else:
            q.normal(q_mu_mu, q_mu_sigma, value=mu_old, name='means')
        return q


class Enc_rws_mu(_EncMu):
    def __init__(self, K, D, num_hidden, num_nss):
        super(self.__class__, self).__init__()
        self._build(D, K, D, num_hidden, num_nss)

    def params(self, ob, K, priors, S=None):
        x = _x_inst(ob)
        S = ob.shape[0] if ob.dim() == 4 else S
        q_mu, q_sigma = self._heads(self.nss1(x).unsqueeze(0), self.nss2(x).unsqueeze(0), priors)   # data only: once per instance
        return q_mu.expand(S, -1, -1, -1).contiguous(), q_sigma.expand(S, -1, -1, -1).contiguous()

    def params_torch(self, ob, K, priors, S=None):
        """same parameters through PyTorch ops only (autograd graph for training)."""
        x = _x_inst(ob)
        S = ob.shape[0] if ob.dim() == 4 else S
        q_mu, q_sigma = self._heads_torch(self.nss1(x).unsqueeze(0), self.nss2(x).unsqueeze(0), priors)
        return q_mu.expand(S, -1, -1, -1), q_sigma.expand(S, -1, -1, -1)

    def forward(self, ob, K, priors, sampled=True, mu_old=None, EPS=1e-8):
        # the reference's forward() is differentiable: with parameters that take part in autograd the Trace is built
        # from the PyTorch-ops parameters (q['means'].log_prob then carries the graph a custom loss needs)
        par = self.params_torch(ob, K, priors) if needs_grad(self) else self.params(ob, K, priors)
        return self._trace(*par, sampled, mu_old)


class Enc_apg_mu(_EncMu):
    def __init__(self, K, D, num_hidden, num_nss):
        super(self.__class__, self).__init__()
        self._build(D + K, K, D, num_hidden, num_nss)

    fused = True          # evaluation path: per-instance label tables + csrc/dmm_fused.cu (z must be one-hot)

    def label_tables(self, x):
        """nss1 / softmax(nss2) of cat(x_n, e_l) for every label l: (B,N,K,J), (B,N,K,K).  z is one-hot, so these are
        ALL the values the per-point MLPs can take -- computed once per instance instead of once per particle."""
        B, N, D = x.shape
        K = self.nss1[0].weight.shape[1] - D
        eye = torch.eye(K, dtype=x.dtype, device=x.device)
        inp = torch.cat((x.unsqueeze(2).expand(B, N, K, D), eye.expand(B, N, K, K)), -1)
        return self.nss1(inp).contiguous(), self.nss2(inp).contiguous()

    def params(self, ob, z, beta, K, priors, tables=None):
        """tables: label_tables(x) computed by the caller (they depend on the data and the weights only, so one
        evaluation serves every sweep of an objective)."""
        x = _x_inst(ob)
        if self.fused:
            # forward values only (the draws carry no gradient); training rebuilds log q through params_torch()
            with torch.no_grad():
                tab1, tab2 = self.label_tables(x) if tables is None else tables
                nss_prior = kernels.nss_pool_onehot(tab1, tab2, z, priors[0], priors[1])
                return self._mu_heads(nss_prior)
        h1 = torch.tanh(_split_linear(self.nss1[0], x, z))
        h2 = torch.tanh(_split_linear(self.nss2[0], x, z))
        return self._heads(self.nss1[2](h1), F.softmax(self.nss2[2](h2), -1), priors)

    def params_torch(self, ob, z, beta, K, priors):
        """same parameters through PyTorch ops only (autograd graph for training)."""
        x = _x_inst(ob)
        h1 = torch.tanh(_split_linear(self.nss1[0], x, z))
        h2 = torch.tanh(_split_linear(self.nss2[0], x, z))
        return self._heads_torch(self.nss1[2](h1), F.softmax(self.nss2[2](h2), -1), priors)

    def forward(self, ob, z, beta, K, priors, sampled=True, mu_old=None, EPS=1e-8):
        # differentiable like the reference's forward() when the parameters take part in autograd (see Enc_rws_mu)
        par = self.params_torch(ob, z, beta, K, priors) if needs_grad(self) else self.params(ob, z, beta, K, priors)
        return self._trace(*par, sampled, mu_old)


class Enc_apg_local(nn.Module):
    def __init__(self, K, D, num_hidden):
        super(self.__class__, self).__init__()
        self.K = K
        self.pi_log_prob = nn.Sequential(nn.Linear(D, num_hidden), nn.ReLU(), nn.Linear(num_hidden, 1))
        self.angle_log_con1 = nn.Sequential(nn.Linear(D, num_hidden), nn.Tanh(), nn.Linear(num_hidden, 1))
        self.angle_log_con0 = nn.Sequential(nn.Linear(D, num_hidden), nn.Tanh(), nn.Linear(num_hidden, 1))

    fused = True          # evaluation path: csrc/dmm_fused.cu

    def can_fuse(self):
        """the fused kernel is compiled for a fixed list of (K, D) (apgs_dmm_supported); other shapes use probs() / cons()"""
        if not self.fused:
            return False
        K, D = getattr(self, 'K', None), self.pi_log_prob[0].weight.shape[1]
        return K is None or bool(kernels.supported(K, D) & 1)

    def sample_fused(self, ob, mu, z_old=None, uniforms=None, z_forced=None, seed=None, decoder=None):
        """probs + categorical draw + Beta concentrations (for z and, when given, for z_old) in one kernel:
        q_probs, z, con1, con0, con1_old, con0_old; with seed also beta ~ Beta(con1, con0) drawn in the kernel and, with
        decoder = Decoder.recon_mu, the decoder's hidden2 of it (kernels.local_encoder_fused)."""
        return kernels.local_encoder_fused(_x_inst(ob), mu, self, uniforms=uniforms, z_forced=z_forced, z_old=z_old,
                                           seed=seed, decoder=decoder)

    def probs(self, ob, mu, K):
        """q_probs (S,B,N,K): MLP on x_n - mu_k (models.py:78-79); W (x - mu) + b = (W x + b) - W mu."""
        x = _x_inst(ob)
        l0 = self.pi_log_prob[0]
        hx = F.linear(x, l0.weight, l0.bias)                       # (B,N,H)
        hm = F.linear(mu, l0.weight)                               # (S,B,K,H)
        h = F.relu(hx[None, :, :, None, :] - hm[:, :, None, :, :])  # (S,B,N,K,H)
        return F.softmax(self.pi_log_prob[2](h).squeeze(-1), -1)

    def cons(self, ob, mu, z):
        """Beta concentrations (S,B,N,1) given z (models.py:86-88, 99-101)."""
        x = _x_inst(ob)
        D = x.shape[-1]
        mu_expand = torch.gather(mu, -2, z.argmax(-1).unsqueeze(-1).expand(-1, -1, -1, D))
        d = x.unsqueeze(0) - mu_expand
        return self.angle_log_con1(d).exp(), self.angle_log_con0(d).exp()

    def forward(self, ob, mu, K, sampled=True, z_old=None, beta_old=None):
        q = Trace()
        q_probs = self.probs(ob, mu, K)
        if sampled:
            z = cat(q_probs, validate_args=False).sample()
            _ = q.variable(cat, probs=q_probs, value=z, name='states')
            c1, c0 = self.cons(ob, mu, z)
            beta = Beta(c1, c0, validate_args=False).sample()
            q.beta(c1, c0, value=beta, name='angles')
        else:
            _ = q.variable(cat, probs=q_probs, value=z_old, name='states')
            c1, c0 = self.cons(ob, mu, z_old)
            q.beta(c1, c0, value=beta_old, name='angles')
        return q


class Decoder(nn.Module):
    def __init__(self, K, D, num_hidden, recon_sigma, CUDA, DEVICE):
        super(self.__class__, self).__init__()
        self.recon_mu = nn.Sequential(nn.Linear(1, num_hidden), nn.Tanh(), nn.Linear(num_hidden, D))
        dev = torch.device(DEVICE) if CUDA else torch.device("cpu")
        self.prior_mu_mu = torch.zeros(D, device=dev)
        self.prior_mu_sigma = torch.ones(D, device=dev) * 10.0
        self.prior_pi = torch.ones(K, device=dev) * (1. / K)
        self.prior_con1 = torch.ones(1, device=dev)
        self.prior_con0 = torch.ones(1, device=dev)
        self.recon_sigma = torch.ones(1, device=dev) * recon_sigma
        self.radi = nn.Parameter(torch.ones(1, device=dev))

    fused = True          # evaluation path: csrc/dmm_fused.cu

    def hidden2(self, beta):
        """unit-normalised decoder output (models.py:202-203)."""
        if self.fused and beta.is_cuda and kernels.supported(2, self.recon_mu[2].weight.shape[0]) & 2:
            return kernels.decoder_hidden2(beta, self.recon_mu)      # forward values; the graph: log_likelihood_torch()
        return self.hidden2_torch(beta)

    def hidden2_torch(self, beta):
        hidden = self.recon_mu(beta * 2 * math.pi)
        return hidden / (hidden ** 2).sum(-1).unsqueeze(-1).sqrt()

    def log_likelihood_torch(self, ob, mu, z, beta):
        """Normal(hidden2 * radi + mu[z], recon_sigma).log_prob(ob) summed over D -> (S,B,N), PyTorch ops only
        (autograd graph of the decoder for loss_theta; models.py:202-209)."""
        x = _x_inst(ob)
        D = x.shape[-1]
        mu_expand = torch.gather(mu, -2, z.argmax(-1).unsqueeze(-1).expand(-1, -1, -1, D))
        recon_mu = self.hidden2_torch(beta) * self.radi + mu_expand
        return Normal(recon_mu, self.recon_sigma, validate_args=False).log_prob(x.unsqueeze(0)).sum(-1)

    def forward(self, ob, mu, z, beta):
        p = Trace()
        if ob.dim() == 3:
            ob = ob.unsqueeze(0).expand(z.shape[0], -1, -1, -1)
        S, B, N, D = ob.shape
        p.normal(self.prior_mu_mu, self.prior_mu_sigma, value=mu, name='means')
        _ = p.variable(cat, probs=self.prior_pi, value=z, name='states')
        p.beta(self.prior_con1, self.prior_con0, value=beta, name='angles')
        mu_expand = torch.gather(mu, -2, z.argmax(-1).unsqueeze(-1).expand(-1, -1, -1, D))
        recon_mu = self.hidden2(beta) * self.radi + mu_expand
        p.normal(recon_mu, self.recon_sigma.repeat(S, B, N, D), value=ob, name='likelihood')
        return p


def init_apg_models(K, D, num_hidden_mu, num_nss, num_hidden_local, num_hidden_dec, recon_sigma, CUDA, device,
                    load_version=None, lr=None):
    """apgs/dmm/apg_training.py:69-101 (same construction order -> same weights under a seed)."""
    enc_rws_mu = Enc_rws_mu(K, D, num_hidden_mu, num_nss)
    enc_apg_local = Enc_apg_local(K, D, num_hidden_local)
    enc_apg_mu = Enc_apg_mu(K, D, num_hidden_mu, num_nss)
    dec = Decoder(K, D, num_hidden_dec, recon_sigma, CUDA, device)
    if CUDA:
        for m in (enc_rws_mu, enc_apg_local, enc_apg_mu, dec):
            m.to(device)
    if load_version is not None:
        weights = torch.load("weights/cp-%s" % load_version)
        enc_rws_mu.load_state_dict(weights['enc-rws-mu'])
        enc_apg_local.load_state_dict(weights['enc-apg-local'])
        enc_apg_mu.load_state_dict(weights['enc-apg-mu'])
        dec.load_state_dict(weights['dec'])
    if lr is not None:
        assert isinstance(lr, float)
        optimizer = torch.optim.Adam(list(enc_rws_mu.parameters()) + list(enc_apg_local.parameters())
                                     + list(enc_apg_mu.parameters()) + list(dec.parameters()), lr=lr, betas=(0.9, 0.99))
        return (enc_rws_mu, enc_apg_local, enc_apg_mu, dec), optimizer
    for m in (enc_rws_mu, enc_apg_local, enc_apg_mu, dec):
        for p in m.parameters():
            p.requires_grad = False
    return (enc_rws_mu, enc_apg_local, enc_apg_mu, dec)


def save_apg_models(models, save_version):
    """apgs/dmm/apg_training.py:103-116: weights/cp-<version> with keys enc-rws-mu / enc-apg-local / enc-apg-mu / dec."""
    from .. import checkpoint
    return checkpoint.save("dmm", models, save_version)

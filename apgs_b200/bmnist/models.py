"""Bouncing-MNIST networks: same classes, constructor arguments, parameter names (state_dict keys) and
forward() contract as apgs/bmnist/models.py (Enc_coor :7-38, Enc_digit :41-74, Dec_coor :77-113,
Dec_digit :116-154).  The MLPs and the random draws stay in PyTorch; the paste -> sum_k -> clamp ->
Bernoulli log-likelihood tail of Dec_digit and the Normal log-densities run in csrc/bmnist.cu.
"""
import math

import torch
import torch.nn as nn
from torch.distributions.normal import Normal

from ..rng import normal_sample
from ..trace import Trace


def needs_grad(*modules):
    """True when any parameter of the given modules takes part in autograd (training mode)."""
    return torch.is_grad_enabled() and any(p.requires_grad for m in modules if isinstance(m, nn.Module)
                                           for p in m.parameters())


class Enc_coor(nn.Module):
    """encoder of the digit positions (models.py:7-38)"""

    def __init__(self, num_pixels, num_hidden, z_where_dim):
        super(self.__class__, self).__init__()
        self.enc_hidden = nn.Sequential(nn.Linear(num_pixels, num_hidden), nn.ReLU())
        self.where_mean = nn.Sequential(nn.Linear(num_hidden, int(0.5 * num_hidden)), nn.ReLU(),
                                        nn.Linear(int(0.5 * num_hidden), z_where_dim), nn.Tanh())
        self.where_log_std = nn.Sequential(nn.Linear(num_hidden, int(0.5 * num_hidden)), nn.ReLU(),
                                           nn.Linear(int(0.5 * num_hidden), z_where_dim))

    fused = True          # evaluation path: the 4761 -> 400 layer on the tensor cores (3xTF32, csrc/linear_tf32x3.cu)
    linear_kernel = "tc5pk"  # "tc5pk": tcgen05 from packed operands (weights packed once, activations written packed by the
    #                          template-match kernel); "tc5": tcgen05, operands staged through registers; "mma": warp-level mma.sync

    def _fusable(self):
        return self.fused and not needs_grad(self) and self.enc_hidden[0].in_features >= 1024

    def wants_packed(self, FP, DP):
        """True when params() takes the template-match map as kernels.PackedRows (objectives._propose asks before matching)."""
        from . import kernels
        lin = self.enc_hidden[0]
        return (self._fusable() and self.linear_kernel == "tc5pk" and lin.weight.is_cuda and lin.weight.dtype == torch.float32
                and kernels.template_match_supports_packed(FP, DP) and lin.in_features == (FP - DP + 1) ** 2)

    def packed_weight(self):
        """the first layer's weight as the packed TF32 operand, rebuilt when the parameter changes (in-place updates bump
        `_version`; a new storage changes `data_ptr`)"""
        from . import kernels
        w = self.enc_hidden[0].weight
        key = (w.data_ptr(), w._version, w.device)
        cache = getattr(self, "_wpack", None)
        if cache is None or cache[0] != key:
            cache = (key, kernels.pack_rows(w.detach(), 1))
            self._wpack = cache
        return cache[1]

    fused_heads = True    # evaluation path: where_mean / where_log_std (+ tanh / exp) in one launch (csrc/bmnist_heads.cu)

    def _heads_fusable(self):
        m, s = self.where_mean, self.where_log_std
        return (self.fused_heads and len(m) == 4 and len(s) == 3 and isinstance(m[0], nn.Linear) and isinstance(m[2], nn.Linear)
                and isinstance(s[0], nn.Linear) and isinstance(s[2], nn.Linear) and isinstance(m[3], nn.Tanh)
                and m[0].weight.shape == s[0].weight.shape and m[2].weight.shape == s[2].weight.shape
                and m[0].bias is not None and s[0].bias is not None and m[2].bias is not None and s[2].bias is not None)

    def heads(self, hidden):
        """(q_mean, q_std) from the hidden layer (models.py:31-32)"""
        if self._fusable() and hidden.is_cuda and hidden.dtype == torch.float32 and self._heads_fusable():
            from . import kernels
            m, s = self.where_mean, self.where_log_std
            ws = (m[0].weight, m[0].bias, s[0].weight, s[0].bias)
            key = tuple((w.data_ptr(), w._version) for w in ws) + (m[0].weight.device,)
            cache = getattr(self, "_hpack", None)
            if cache is None or cache[0] != key:
                cache = (key, kernels.coor_heads_weights(*ws))
                self._hpack = cache
            w1t, b1 = cache[1]
            return kernels.coor_heads(hidden, w1t, b1, m[2].weight, m[2].bias, s[2].weight, s[2].bias)
        return self.where_mean(hidden), self.where_log_std(hidden).exp()

    def params(self, conved):
        from . import kernels
        lin = self.enc_hidden[0]
        if isinstance(conved, kernels.PackedRows):
            if self._fusable() and self.linear_kernel == "tc5pk":
                return self.heads(kernels.linear_tc5_packed(conved, self.packed_weight(), lin.bias, relu=True))
            conved = conved.unpack()
        if self._fusable() and conved.is_cuda and conved.dtype == torch.float32:
            fn = kernels.linear_tf32x3 if self.linear_kernel == "mma" else kernels.linear_tc5
            hidden = fn(conved, lin.weight, lin.bias, relu=True)                          # Linear + ReLU, fp32-level accuracy
        else:
            hidden = self.enc_hidden(conved)
        return self.heads(hidden)

    def forward(self, conved, sampled=True, z_where_old=None):
        q = Trace()
        q_mean, q_std = self.params(conved)
        if sampled:
            z_where = normal_sample(q_mean, q_std)
            q.normal(loc=q_mean, scale=q_std, value=z_where, name='z_where')
        else:
            q.normal(loc=q_mean, scale=q_std, value=z_where_old, name='z_where')
        return q


class Enc_digit(nn.Module):
    """encoder of digit features (models.py:41-74)"""

    def __init__(self, num_pixels, num_hidden, z_what_dim):
        super(self.__class__, self).__init__()
        self.enc_hidden = nn.Sequential(nn.Linear(num_pixels, num_hidden), nn.ReLU(),
                                        nn.Linear(num_hidden, int(0.5 * num_hidden)), nn.ReLU())
        self.q_mean = nn.Sequential(nn.Linear(int(0.5 * num_hidden), z_what_dim))
        self.q_log_std = nn.Sequential(nn.Linear(int(0.5 * num_hidden), z_what_dim))

    def params(self, cropped):
        hidden = self.enc_hidden(cropped).mean(2)          # T is the 3rd dim of cropped
        return self.q_mean(hidden), self.q_log_std(hidden).exp()

    def forward(self, cropped, sampled=True, z_what_old=None):
        q = Trace()
        q_mu, q_std = self.params(cropped)
        if sampled:
            z_what = normal_sample(q_mu, q_std)
            q.normal(loc=q_mu, scale=q_std, value=z_what, name='z_what')
        else:
            q.normal(loc=q_mu, scale=q_std, value=z_what_old, name='z_what')
        return q


class Dec_coor():
    """generative model of digit positions (models.py:77-113): z_1 ~ N(0, 1), z_t | z_t-1 ~ N(z_t-1, 0.2)"""

    def __init__(self, z_where_dim, CUDA, device):
        dev = torch.device(device) if CUDA else torch.device("cpu")
        self.prior_mu0 = torch.zeros(z_where_dim, device=dev)
        self.prior_Sigma0 = torch.ones(z_where_dim, device=dev) * 1.0
        self.prior_Sigmat = torch.ones(z_where_dim, device=dev) * 0.2

    def forward(self, z_where_t, z_where_t_1=None, disp=None):
        if z_where_t_1 is None:
            return Normal(self.prior_mu0, self.prior_Sigma0).log_prob(z_where_t).sum(-1)
        return Normal(z_where_t_1, self.prior_Sigmat).log_prob(z_where_t).sum(-1)

    def log_prior(self, z_where_t, z_where_t_1=None, disp=None):
        if z_where_t_1 is None:
            return Normal(self.prior_mu0, self.prior_Sigma0).log_prob(z_where_t).sum(-1).sum(-1)
        return Normal(z_where_t_1, self.prior_Sigmat).log_prob(z_where_t).sum(-1).sum(-1)


class Dec_digit(nn.Module):
    """decoder of the digit features (models.py:116-154)"""

    def __init__(self, num_pixels, num_hidden, z_what_dim, CUDA, device):
        super(self.__class__, self).__init__()
        self.digit_mean = nn.Sequential(nn.Linear(z_what_dim, int(0.5 * num_hidden)), nn.ReLU(),
                                        nn.Linear(int(0.5 * num_hidden), num_hidden), nn.ReLU(),
                                        nn.Linear(num_hidden, num_pixels), nn.Sigmoid())
        dev = torch.device(device) if CUDA else torch.device("cpu")
        self.prior_mu = torch.zeros(z_what_dim, device=dev)
        self.prior_std = torch.ones(z_what_dim, device=dev)

    def templates(self, z_what):
        digit_mean = self.digit_mean(z_what)
        S, B, K, DP2 = digit_mean.shape
        DP = int(math.sqrt(DP2))
        return digit_mean.view(S, B, K, DP, DP)

    def forward(self, frames, z_what, z_where=None, AT=None):
        digit_mean = self.templates(z_what)
        if z_where is None:
            return digit_mean.detach()
        assert AT is not None, "ERROR! NoneType variable AT found."
        assert frames is not None, "ERROR! NoneType variable frames found."
        # fused: paste, sum over digits, clamp, Bernoulli log-likelihood (csrc/bmnist.cu k_stn_paste<true>)
        ll, recon_frames = AT.reconstruct_ll(digit_mean, z_where, frames, want_recon=True)
        log_prior = Normal(loc=self.prior_mu, scale=self.prior_std).log_prob(z_what).sum(-1)
        return log_prior, ll, recon_frames


def init_models(frame_pixels, digit_pixels, num_hidden_digit, num_hidden_coor, z_where_dim, z_what_dim, CUDA, device,
                load_version=None, lr=None):
    """apgs/bmnist/apg_training.py:62-92 (same construction order -> same weights under a seed)."""
    enc_coor = Enc_coor(num_pixels=(frame_pixels - digit_pixels + 1) ** 2, num_hidden=num_hidden_coor, z_where_dim=z_where_dim)
    dec_coor = Dec_coor(z_where_dim=z_where_dim, CUDA=CUDA, device=device)
    enc_digit = Enc_digit(num_pixels=digit_pixels ** 2, num_hidden=num_hidden_digit, z_what_dim=z_what_dim)
    dec_digit = Dec_digit(num_pixels=digit_pixels ** 2, num_hidden=num_hidden_digit, z_what_dim=z_what_dim, CUDA=CUDA, device=device)
    if CUDA:
        enc_coor.to(device)
        enc_digit.to(device)
        dec_digit.to(device)
    if load_version is not None:
        weights = torch.load("weights/cp-%s" % load_version)
        enc_coor.load_state_dict(weights['enc-coor'])
        enc_digit.load_state_dict(weights['enc-digit'])
        dec_digit.load_state_dict(weights['dec-digit'])
    if lr is not None:
        optimizer = torch.optim.Adam(list(enc_coor.parameters()) + list(enc_digit.parameters()) + list(dec_digit.parameters()),
                                     lr=lr, betas=(0.9, 0.99))
        return (enc_coor, dec_coor, enc_digit, dec_digit), optimizer
    for m in (enc_coor, enc_digit, dec_digit):
        for p in m.parameters():
            p.requires_grad = False
    return (enc_coor, dec_coor, enc_digit, dec_digit)


def save_models(models, save_version):
    """apgs/bmnist/apg_training.py:95-104: weights/cp-<version> with keys enc-coor / enc-digit / dec-digit."""
    from .. import checkpoint
    return checkpoint.save("bmnist", models, save_version)

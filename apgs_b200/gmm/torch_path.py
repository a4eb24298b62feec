"""Generic-shape path of the GMM reweighting arithmetic in stock PyTorch CUDA ops.

The sm_100a kernels are compiled for a fixed list of (K, D) (csrc/gmm.cu APGS_KD_DISPATCH; `apgs_gmm_supported`); the
reference accepts any --num_clusters / data dimension.  For shapes outside that list the host wrappers in kernels.py
route here instead of failing: the same quantities (G1-G4 of SURVEY.md section 8a) computed with torch.distributions on the
tensors' own (CUDA) device -- library calls, slower, numerically the reference's own formulas
(apgs/gmm/kls_gmm.py:92-119, apgs/gmm/models.py:139-185, apgs/gmm/objectives.py:74-227).  Not a CPU fallback: CPU
tensors are rejected exactly as on the kernel path.
"""
import torch
from torch.distributions.gamma import Gamma
from torch.distributions.normal import Normal


class Weights(object):
    __slots__ = ("log_w", "log_q_f", "density", "log_w_f", "log_w_b", "log_w_local", "log_q_local")

    def __init__(self):
        for n in self.__slots__:
            setattr(self, n, None)


def _x4(x, S):
    return x if x.dim() == 4 else x.unsqueeze(0).expand(S, -1, -1, -1)


def nss_posterior(xp, gamma, prior):
    """kls_gmm.py:92-119 incl. quirks Q1 (zero counts replaced everywhere) and Q2 ((x_bar - prior_nu)^2)."""
    s1 = gamma.sum(2).unsqueeze(-1)
    s2 = torch.einsum('sbnk,sbnd->sbkd', gamma, xp)
    s3 = torch.einsum('sbnk,sbnd->sbkd', gamma, xp * xp)
    cnt = s1.expand_as(s2)
    cnt = torch.where(cnt == 0.0, torch.ones_like(cnt), cnt)
    x_bar = s2 / cnt
    alpha = prior.prior_alpha + cnt / 2
    nu = prior.prior_nu + cnt
    mu = (prior.prior_mu * prior.prior_nu + s2) / (cnt + prior.prior_nu)
    beta = prior.prior_beta + (s3 - (s2 ** 2) / cnt) / 2. + (cnt * prior.prior_nu / (cnt + prior.prior_nu)) * ((x_bar - prior.prior_nu) ** 2) / 2.
    return alpha, beta, mu, nu


def _ng_lp(tau, mu, a, b, m, n):
    lp = Gamma(a, b, validate_args=False).log_prob(tau) + Normal(m, 1. / (n * tau).sqrt(), validate_args=False).log_prob(mu)
    return lp.sum(-1).sum(-1)


def _ll(x, z, tau, mu):
    """log N(x_n; mu_{z_n}, tau_{z_n}^-1/2) summed over D -> (S,B,N)"""
    lab = z.argmax(-1).unsqueeze(-1).expand(-1, -1, -1, x.shape[-1])
    return Normal(torch.gather(mu, 2, lab), torch.gather(1. / tau.sqrt(), 2, lab), validate_args=False).log_prob(x).sum(-1)


def _cat_lp(z, probs):
    """OneHotCategorical(probs).log_prob(z) with torch's eps clamp (quirk Q4) -> (S,B,N)"""
    p = probs / probs.sum(-1, keepdim=True)
    eps = torch.finfo(p.dtype).eps
    return torch.log(p.clamp(eps, 1 - eps)).gather(-1, z.argmax(-1, keepdim=True)).squeeze(-1)


def _prior_ng(prior):
    return prior.prior_alpha, prior.prior_beta, prior.prior_mu, prior.prior_nu


def logw_oneshot(x, post, prior, tau, mu, q_probs, z, parts=False):
    x = _x4(x, z.shape[0])
    w = Weights()
    w.log_q_f = _ng_lp(tau, mu, *post) + _cat_lp(z, q_probs).sum(-1)
    w.density = _ll(x, z, tau, mu).sum(-1) + _ng_lp(tau, mu, *_prior_ng(prior)) + _cat_lp(z, prior.prior_pi.expand_as(q_probs)).sum(-1)
    w.log_w = w.density - w.log_q_f
    if parts:
        w.log_w_f = w.log_w
    return w


def logw_eta(x, z, post, prior, tau, mu, tau_old, mu_old, parts=False):
    x = _x4(x, z.shape[0])
    w = Weights()
    w.log_q_f = _ng_lp(tau, mu, *post)
    w.density = _ng_lp(tau, mu, *_prior_ng(prior))
    lwf = _ll(x, z, tau, mu).sum(-1) + w.density - w.log_q_f
    lwb = _ll(x, z, tau_old, mu_old).sum(-1) + _ng_lp(tau_old, mu_old, *_prior_ng(prior)) - _ng_lp(tau_old, mu_old, *post)
    w.log_w = lwf - lwb
    if parts:
        w.log_w_f, w.log_w_b = lwf, lwb
    return w


def logw_z(x, tau, mu, q_probs, prior, z, z_old, parts=False, local=False):
    x = _x4(x, z.shape[0])
    w = Weights()
    pi = prior.prior_pi.expand_as(q_probs)
    lq_f, lq_b = _cat_lp(z, q_probs), _cat_lp(z_old, q_probs)
    lp_f = _ll(x, z, tau, mu) + _cat_lp(z, pi)
    lp_b = _ll(x, z_old, tau, mu) + _cat_lp(z_old, pi)
    loc = (lp_f - lq_f) - (lp_b - lq_b)
    w.log_w, w.log_q_f, w.density = loc.sum(-1), lq_f.sum(-1), lp_f.sum(-1)
    if parts:
        w.log_w_f, w.log_w_b = (lp_f - lq_f).sum(-1), (lp_b - lq_b).sum(-1)
    if local:
        w.log_w_local, w.log_q_local = loc, lq_f
    return w


def logw_joint(x, post, prior, tau, mu, tau_old, mu_old, q_probs, z, z_old, parts=False):
    x = _x4(x, z.shape[0])
    w = Weights()
    pi = prior.prior_pi.expand_as(q_probs)
    w.log_q_f = _ng_lp(tau, mu, *post) + _cat_lp(z, q_probs).sum(-1)
    w.density = _ll(x, z, tau, mu).sum(-1) + _ng_lp(tau, mu, *_prior_ng(prior)) + _cat_lp(z, pi).sum(-1)
    lwf = w.density - w.log_q_f
    lwb = (_ll(x, z_old, tau_old, mu_old).sum(-1) + _ng_lp(tau_old, mu_old, *_prior_ng(prior)) + _cat_lp(z_old, pi).sum(-1)
           - _ng_lp(tau_old, mu_old, *post) - _cat_lp(z_old, q_probs).sum(-1))
    w.log_w = lwf - lwb
    if parts:
        w.log_w_f, w.log_w_b = lwf, lwb
    return w

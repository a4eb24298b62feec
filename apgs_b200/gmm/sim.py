"""Synthetic GMM instances with the reference simulator's distributions (apgs/gmm/sim_gmm.py:30-38,
priors :78-81): tau ~ Gamma(2,2), mu ~ N(0, (0.1 tau)^-1/2), z ~ Cat(1/K)^N, x ~ N(mu_z, tau_z^-1/2)."""
import torch
from torch.distributions.gamma import Gamma
from torch.distributions.normal import Normal


def simulate(B, N=60, K=3, D=2, alpha=2.0, beta=2.0, mu0=0.0, nu0=0.1):
    """returns x (B, N, D) on the CPU (callers move / pin it)."""
    tau = Gamma(torch.ones(B, K, D) * alpha, torch.ones(B, K, D) * beta).sample()
    mean = Normal(torch.ones(B, K, D) * mu0, 1. / (tau * nu0).sqrt()).sample()
    labels = torch.randint(0, K, (B, N))
    idx = labels.unsqueeze(-1).expand(B, N, D)
    return Normal(torch.gather(mean, 1, idx), torch.gather(1. / tau.sqrt(), 1, idx)).sample()

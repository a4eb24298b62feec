"""Synthetic GMM instances with the reference simulator's distributions (apgs/gmm/sim_gmm.py:30-38,
priors :78-81): tau ~ Gamma(2,2), mu ~ N(0, (0.1 tau)^-1/2), z ~ Cat(1/K)^N, x ~ N(mu_z, tau_z^-1/2)."""
import os

import numpy as np
import torch
from torch.distributions.gamma import Gamma
from torch.distributions.normal import Normal


def simulate(B, N=60, K=3, D=2, alpha=2.0, beta=2.0, mu0=0.0, nu0=0.1, return_latents=False):
    """returns x (B, N, D) on the CPU (callers move / pin it); with return_latents also the one-hot assignments
    (B,N,K), the precisions and the means (B,K,D)."""
    tau = Gamma(torch.ones(B, K, D) * alpha, torch.ones(B, K, D) * beta).sample()
    mean = Normal(torch.ones(B, K, D) * mu0, 1. / (tau * nu0).sqrt()).sample()
    labels = torch.randint(0, K, (B, N))
    idx = labels.unsqueeze(-1).expand(B, N, D)
    x = Normal(torch.gather(mean, 1, idx), torch.gather(1. / tau.sqrt(), 1, idx)).sample()
    if return_latents:
        return x, torch.nn.functional.one_hot(labels, K).float(), tau, mean
    return x


def save_dataset(path, num_instances, **kw):
    """the reference's on-disk format (sim_gmm.py:56-68): `<path>/ob.npy` (num, N, D) and `<path>/assignment.npy`
    (num, N, K), float64."""
    os.makedirs(path, exist_ok=True)
    x, z, _, _ = simulate(num_instances, return_latents=True, **kw)
    np.save(os.path.join(path, "ob"), x.double().numpy())
    np.save(os.path.join(path, "assignment"), z.double().numpy())
    return os.path.join(path, "ob.npy"), os.path.join(path, "assignment.npy")


def load_dataset(path):
    """what apgs/gmm/apg_training.py:185-186 does: (data, assignments) as float tensors"""
    return (torch.from_numpy(np.load(os.path.join(path, "ob.npy"))).float(),
            torch.from_numpy(np.load(os.path.join(path, "assignment.npy"))).float())

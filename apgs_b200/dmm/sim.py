"""Synthetic DMM (ring mixture) instances with the reference simulator's distributions (apgs/dmm/sim_dmm.py:27-62,
defaults :93-99): K rings of radius 2 whose centres are N(0, 3^2); each ring holds N/K points laid out as `period`
interleaved circles of N/(K period) equally spaced angles in [0, 2 pi], every circle shifted by its own U(0, spacing)
offset; isotropic N(0, 0.1^2) noise.  Vectorised over the B instances (the reference loops in numpy)."""
import math
import os

import numpy as np
import torch


def simulate(B, N=200, K=4, D=2, period=2, radius=2.0, noise=0.1, center_std=3.0, return_latents=False):
    """x (B, N, D) on the CPU; with return_latents also (one-hot states (B,N,K), centres (B,K,D), angles (B,N,1))."""
    if D != 2:
        raise ValueError("rings live in the plane (D = 2)")
    Nk = N // K
    per = Nk // period
    if Nk * K != N or per * period != Nk or per < 2:
        raise ValueError("N must be a multiple of K * period (N=%d, K=%d, period=%d)" % (N, K, period))
    base = torch.linspace(0, 2 * math.pi, per)                                    # endpoint included, as np.linspace
    spacing = float(base[1] - base[0])
    shift = torch.rand(B, K, period, 1) * spacing
    ang = (base.view(1, 1, 1, per) + shift).reshape(B, K, Nk)
    ring = torch.stack((torch.cos(ang), torch.sin(ang)), -1) * radius             # B,K,Nk,2
    centres = torch.randn(B, K, 1, D) * center_std
    x = (ring + torch.randn(B, K, Nk, D) * noise + centres).reshape(B, N, D)
    if not return_latents:
        return x
    states = torch.eye(K).view(1, K, 1, K).expand(B, K, Nk, K).reshape(B, N, K).contiguous()
    return x, states, centres.squeeze(2), ang.reshape(B, N, 1)


def save_dataset(path, num_instances, **kw):
    """the reference's on-disk format: `<path>/ob.npy`, float64 (num_instances, N, D) (sim_dmm.py:77-86)."""
    os.makedirs(path, exist_ok=True)
    ob = simulate(num_instances, **kw).double().numpy()
    np.save(os.path.join(path, "ob"), ob)
    return os.path.join(path, "ob.npy")


def load_dataset(path):
    """what apgs/dmm/apg_training.py:185 does: torch.from_numpy(np.load(path + 'ob.npy')).float()"""
    f = path if path.endswith(".npy") else os.path.join(path, "ob.npy")
    return torch.from_numpy(np.load(f)).float()

"""Synthetic bouncing-digit sequences with the reference simulator's dynamics (apgs/bmnist/sim_bmnist.py:48-114):
start x0 ~ U(-1,1)^2, unit velocity in a U(0, 2 pi) direction, X_{t+1} = X_t + V dt (dt = 0.3) with elastic reflection
at +-1, digits pasted by the spatial transformer (scale FP/DP, translation x (FP-DP)/DP, nearest, align_corners) and
the K digit layers summed and clamped to [0,1].  MNIST itself cannot be downloaded here (sim_bmnist.py:33-46 fetches
it): `templates` takes the digit images ((K,28,28) or (B,K,28,28) in [0,1]); without them smooth synthetic blobs are
used.  On a CUDA device the frames are rendered by this package's own paste kernel (Affine_Transformer.digit_to_frame);
on the CPU by the same separable nearest-index formula in PyTorch ops."""
import math
import os

import numpy as np
import torch


def blobs(K=3, DP=28):
    """K smooth anisotropic blobs in [0,1] standing in for MNIST digits."""
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, DP), torch.linspace(-1, 1, DP), indexing='ij')
    out = []
    for k in range(K):
        a, b = 0.35 + 0.1 * (k % 3), 0.55 - 0.1 * (k % 3)
        out.append(torch.exp(-((xx / a) ** 2 + (yy / b) ** 2) * 2.0).clamp(0, 1))
    return torch.stack(out, 0)


def trajectories(B, T=10, K=3, dt=0.3):
    """positions (B,T,K,2) and velocities (B,T,K,2) of sim_trajectory (sim_bmnist.py:48-77), vectorised."""
    pos = torch.rand(B, K, 2) * 2 - 1
    ang = torch.rand(B, K) * 2 * math.pi
    vel = torch.stack((torch.cos(ang), torch.sin(ang)), -1)
    X, V = [], []
    for _ in range(T):
        X.append(pos)
        V.append(vel)
        nxt = pos + vel * dt
        lo, hi = nxt < -1.0, nxt > 1.0
        nxt = torch.where(lo, -1.0 + (-1.0 - nxt).abs(), nxt)
        nxt = torch.where(hi, 1.0 - (nxt - 1.0).abs(), nxt)
        vel = torch.where(lo | hi, -vel, vel)
        pos = nxt
    return torch.stack(X, 1), torch.stack(V, 1)


def _paste_cpu(digit, z, FP, DP):
    """digit (B,K,DP,DP), z (B,T,K,2) -> (B,T,K,FP,FP): the separable nearest gather of the transformer."""
    lin = torch.linspace(-1, 1, FP)
    scale = float(torch.tensor(FP / DP, dtype=torch.float32))
    tr = (FP - DP) / DP

    def axis(t):
        idx = torch.round(((lin * scale + t.unsqueeze(-1)) + 1) * ((DP - 1) / 2))
        return idx.clamp(0, DP - 1).long(), (idx >= 0) & (idx <= DP - 1)
    col, cok = axis(-1 * (z[..., 0] * tr))
    row, rok = axis(z[..., 1] * tr)
    T = z.shape[1]
    img = digit.unsqueeze(1).expand(-1, T, -1, -1, -1)
    rows = torch.gather(img, -2, row.unsqueeze(-1).expand(row.shape + (DP,)))
    out = torch.gather(rows, -1, col.unsqueeze(-2).expand(col.shape[:-1] + (FP, FP)))
    return out * (rok.unsqueeze(-1) & cok.unsqueeze(-2)).to(out.dtype)


def simulate(B, T=10, K=3, FP=96, DP=28, dt=0.3, templates=None, device=None, return_latents=False):
    """frames (B,T,FP,FP) in [0,1] (on `device` when given, else CPU)."""
    tmpl = blobs(K, DP) if templates is None else templates.float()
    if tmpl.dim() == 3:
        tmpl = tmpl.unsqueeze(0).expand(B, -1, -1, -1)
    X, V = trajectories(B, T, K, dt)
    if device is not None and torch.device(device).type == "cuda":
        from .affine_transformer import Affine_Transformer
        AT = Affine_Transformer(FP, DP, True, device)
        pasted = AT.digit_to_frame(tmpl.contiguous().to(device).unsqueeze(0), X.to(device).unsqueeze(0))[0]
    else:
        pasted = _paste_cpu(tmpl, X, FP, DP)
    frames = pasted.sum(2).clamp(0.0, 1.0)
    return (frames, X, V) if return_latents else frames


def save_dataset(path, num_instances, chunk_size=1000, **kw):
    """the reference's on-disk format: chunks `<path>/ob-<i>.npy` of (<= chunk_size, T, FP, FP) float32
    (sim_bmnist.py:116-156)."""
    os.makedirs(path, exist_ok=True)
    files, left, i = [], num_instances, 1
    while left > 0:
        n = min(chunk_size, left)
        np.save(os.path.join(path, "ob-%d" % i), simulate(n, **kw).cpu().numpy())
        files.append(os.path.join(path, "ob-%d.npy" % i))
        left -= n
        i += 1
    return files


def load_dataset(file):
    """what apgs/bmnist/apg_training.py:21 does per chunk"""
    return torch.from_numpy(np.load(file)).float()

"""Oracle: bmnist spatial transformer (nearest, align_corners=True) in closed form.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Restates apgs/bmnist/affine_transformer.py:41-65 (affine_grid + grid_sample(mode='nearest',
align_corners=True)) for the purely diagonal theta the reference builds.  Because theta is diagonal the
sampling grid is separable:

    src_col[j] = rint( ((lin_W[j] * scale + t_x) + 1) * (W_in - 1)/2 ),   zero outside [0, W_in-1]
    src_row[i] = rint( ((lin_H[i] * scale + t_y) + 1) * (H_in - 1)/2 )

all in fp32 with one rounding per operation; lin = torch.linspace(-1, 1, n) (the affine_grid base grid).
  digit_to_frame (:41-52): scale = FP/DP,  t_x = -z_x * (FP-DP)/DP,  t_y = +z_y * (FP-DP)/DP
  frame_to_digit (:54-65): scale = DP/FP,  t_x = +z_x * (FP-DP)/FP,  t_y = -z_y * (FP-DP)/FP
Verified bit-exact against the reference in tests/golden/make_golden.py (section stn).
"""
import numpy as np
import torch

f32 = np.float32


def base_axis(n):
    """affine_grid's base coordinates for align_corners=True."""
    return torch.linspace(-1, 1, n).numpy().astype(f32)


def axis_index(lin, scale, t, n_in):
    """lin (n_out,), t (...,) -> int index (..., n_out) with -1 where the sample falls outside."""
    t = np.asarray(t, dtype=f32)[..., None]
    v = (lin.astype(f32) * f32(scale)).astype(f32)
    v = (v + t).astype(f32)
    u = ((v + f32(1)).astype(f32) * f32((n_in - 1) / 2)).astype(f32)
    idx = np.rint(u)
    ok = (idx >= 0) & (idx <= n_in - 1)
    return np.where(ok, idx, -1).astype(np.int64)


def digit_to_frame(digit, z_where, FP=96, DP=28):
    """digit (S,B,K,DP,DP), z_where (S,B,T,K,2) -> (S,B,T,K,FP,FP)   (paste, zero padded)."""
    digit = np.asarray(digit, dtype=f32)
    z = np.asarray(z_where, dtype=f32)
    S, B, T, K, _ = z.shape
    scale = f32(FP / DP)
    tr = f32((FP - DP) / DP)
    lin = base_axis(FP)
    col = axis_index(lin, scale, -(z[..., 0] * tr).astype(f32), DP)        # (S,B,T,K,FP)
    row = axis_index(lin, scale, (z[..., 1] * tr).astype(f32), DP)
    out = np.zeros((S, B, T, K, FP, FP), dtype=f32)
    for s in range(S):
        for b in range(B):
            for t in range(T):
                for k in range(K):
                    r, c = row[s, b, t, k], col[s, b, t, k]
                    patch = digit[s, b, k][np.clip(r, 0, DP - 1)[:, None], np.clip(c, 0, DP - 1)[None, :]]
                    out[s, b, t, k] = np.where((r >= 0)[:, None] & (c >= 0)[None, :], patch, f32(0))
    return out


def frame_to_digit(frames, z_where, FP=96, DP=28):
    """frames (S,B,T,FP,FP), z_where (S,B,T,K,2) -> (S,B,T,K,DP,DP)   (crop, zero padded)."""
    frames = np.asarray(frames, dtype=f32)
    z = np.asarray(z_where, dtype=f32)
    S, B, T, K, _ = z.shape
    scale = f32(DP / FP)
    tr = f32((FP - DP) / FP)
    lin = base_axis(DP)
    col = axis_index(lin, scale, (z[..., 0] * tr).astype(f32), FP)
    row = axis_index(lin, scale, -(z[..., 1] * tr).astype(f32), FP)
    out = np.zeros((S, B, T, K, DP, DP), dtype=f32)
    for s in range(S):
        for b in range(B):
            for t in range(T):
                for k in range(K):
                    r, c = row[s, b, t, k], col[s, b, t, k]
                    patch = frames[s, b, t][np.clip(r, 0, FP - 1)[:, None], np.clip(c, 0, FP - 1)[None, :]]
                    out[s, b, t, k] = np.where((r >= 0)[:, None] & (c >= 0)[None, :], patch, f32(0))
    return out


def recon_and_ll(digit, z_where, frames, FP=96, DP=28, EPS=1e-9):
    """apgs/bmnist/models.py:148-162: recon = clamp(sum_k paste_k, 0, 1); Bernoulli-like pixel log-likelihood
    summed over the frame -> recon (S,B,T,FP,FP), ll (S,B,T)   (torch-CPU fp32 ops, reference order)."""
    pasted = torch.from_numpy(digit_to_frame(digit, z_where, FP, DP))
    recon = torch.clamp(pasted.sum(-3), min=0.0, max=1.0)
    x = torch.as_tensor(np.asarray(frames, dtype=f32))
    ll = (torch.log(recon + EPS) * x + torch.log(1 - recon + EPS) * (1 - x)).sum(-1).sum(-1)
    return recon, ll

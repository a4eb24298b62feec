"""Oracle: ancestor sampling + ancestor gather (restates apgs/resampler.py:22-54).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Closed forms below were verified bit-exact against the reference run on CPU
(tests/golden/make_golden.py):

  systematic (apgs/resampler.py:26-37)
      pos[b, s] = fl32(fl32(u[b] + s) / S)                      (:28)
      w         = softmax(log_w, dim=0)                         (:30, ATen)
      cdf[b, i] = fl32( sum_{j<=i} (double) w[j, b] )           (:31, ATen CPU cumsum
                                                                 accumulates in double)
      ncdf      = cdf / cdf[:, -1]   (max of a monotone row)    (:32-33)
      a[s, b]   = #{ i : ncdf[b, i] < pos[b, s] }               (:35, searchsorted left)

  multinomial (apgs/resampler.py:39-41 -> Categorical(w.T).sample((S,)))
      p         = w.T / w.T.sum(-1)     (Categorical.__init__ renormalises, categorical.py:70;
                                         `categorical_probs` below -- the Level-1 input)
      then torch.multinomial(p, S, True) on ATen CPU:
      cdf[b, i] = sequential fp32 running sum of p[:, b]
      ncdf      = cdf / cdf[:, -1]                              (fp32)
      a[s, b]   = #{ i : (double) ncdf[b, i] < u[b, s] },  u double in [0, 1)
      (u = torch.rand(B, S, dtype=float64) reproduces the generator stream of the
       reference call exactly.)

  gather (apgs/resampler.py:46-54)
      out[s, b, ...] = var[a[s, b], b, ...]
"""
import numpy as np
import torch


def softmax_weights(log_w: torch.Tensor) -> torch.Tensor:
    """apgs/resampler.py:30 -- normalised weights over the particle axis (dim 0)."""
    return torch.softmax(log_w, 0)


def weight_stats(log_w: torch.Tensor):
    """W1: w, ess, lse.  apgs/gmm/objectives.py:93,98 (ess = 1 / sum_s w^2); lse is the
    (implicit) log-marginal term torch.logsumexp(log_w, 0)."""
    w = torch.softmax(log_w, 0)
    ess = 1.0 / (w ** 2).sum(0)
    lse = torch.logsumexp(log_w, 0)
    return w, ess, lse


def local_weight_stats(log_w_local: torch.Tensor):
    """W2: per-point softmax over S and the summed resampling weight.
    apgs/gmm/objectives.py:216-218, apgs/dmm/objectives.py:173-175."""
    w_local = torch.softmax(log_w_local, 0)
    log_w = log_w_local.sum(-1)
    return w_local, log_w


def categorical_probs(w: torch.Tensor) -> torch.Tensor:
    """torch/distributions/categorical.py:70 applied to the transposed weights
    (apgs/resampler.py:41): (S,B) softmax weights -> (S,B) probabilities handed to
    torch.multinomial.  Uses ATen's own reduction, like the reference."""
    wt = w.transpose(0, 1)
    return (wt / wt.sum(-1, keepdim=True)).transpose(0, 1).contiguous()


def systematic_positions(u: np.ndarray, S: int) -> np.ndarray:
    """apgs/resampler.py:28 -- (B,S) fp32 strata positions, one uniform per instance."""
    u = np.asarray(u, dtype=np.float32).reshape(-1, 1)
    spacing = np.arange(S, dtype=np.float32)[None, :]
    return ((u + spacing).astype(np.float32) / np.float32(S)).astype(np.float32)


def cdf_f64acc(w: np.ndarray) -> np.ndarray:
    """(S,B) fp32 weights -> (B,S) fp32 normalised CDF with double accumulation
    (apgs/resampler.py:31-33 as executed by ATen on CPU)."""
    c = np.cumsum(np.ascontiguousarray(w.T).astype(np.float64), axis=1).astype(np.float32)
    return (c / c[:, -1:]).astype(np.float32)


def cdf_seq_f32(w: np.ndarray) -> np.ndarray:
    """(S,B) fp32 weights -> (B,S) fp32 normalised CDF with a sequential fp32 running
    sum (ATen CPU multinomial_with_replacement)."""
    wt = np.ascontiguousarray(w.T).astype(np.float32)
    c = np.empty_like(wt)
    acc = np.zeros(wt.shape[0], dtype=np.float32)
    for j in range(wt.shape[1]):
        acc = (acc + wt[:, j]).astype(np.float32)
        c[:, j] = acc
    return (c / c[:, -1:]).astype(np.float32)


def systematic_ancestors(w, u) -> np.ndarray:
    """Level-1 contract: normalised weights (S,B) fp32 + uniforms (B,) fp32 -> int64 (S,B)."""
    w = np.asarray(w, dtype=np.float32)
    S, B = w.shape
    ncdf = cdf_f64acc(w)
    pos = systematic_positions(u, S)
    a = np.empty((B, S), dtype=np.int64)
    for b in range(B):
        a[b] = np.searchsorted(ncdf[b], pos[b], side="left")
    return np.ascontiguousarray(a.T)


def multinomial_ancestors(w, u, cdf_mode: str = "seq_f32") -> np.ndarray:
    """Level-1 contract: probabilities (S,B) fp32 as handed to torch.multinomial (see
    categorical_probs) + uniforms (B,S) float64 -> int64 (S,B).
    cdf_mode 'seq_f32' is the ATen-CPU semantic; 'f64' is the double-accumulated CDF the
    CUDA path switches to for very long particle axes (documented deviation)."""
    w = np.asarray(w, dtype=np.float32)
    S, B = w.shape
    u = np.asarray(u, dtype=np.float64).reshape(B, -1)
    ncdf = (cdf_seq_f32(w) if cdf_mode == "seq_f32" else cdf_f64acc(w)).astype(np.float64)
    a = np.empty((B, u.shape[1]), dtype=np.int64)
    for b in range(B):
        a[b] = np.searchsorted(ncdf[b], u[b], side="left")
    return np.ascontiguousarray(a.T)


def gather_ancestors(var, a) -> np.ndarray:
    """apgs/resampler.py:46-54: out[s,b,...] = var[a[s,b], b, ...] (any trailing dims)."""
    var = np.asarray(var)
    a = np.asarray(a)
    S, B = a.shape
    cols = np.arange(B)[None, :]
    return var[a, cols]


def resample(latents, log_w, u, strategy="systematic"):
    """apgs/*/objectives.py resample_variables: one ancestor draw, gather every latent."""
    w = softmax_weights(torch.as_tensor(log_w)).numpy()
    if strategy == "systematic":
        a = systematic_ancestors(w, u)
    else:
        a = multinomial_ancestors(categorical_probs(torch.from_numpy(w)).numpy(), u)
    return a, [gather_ancestors(np.asarray(v), a) for v in latents]

"""Oracle: log-densities used on the path, restated from torch.distributions 2.11.0
(the third-party dependency that carries the arithmetic).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  torch-CPU fp32 ops, written in the
same operation order as the library so that rounding matches the reference run on CPU.
"""
import math
import torch

LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


def normal_log_prob(value, loc, scale):
    """torch/distributions/normal.py:87-101:
    -((v-loc)^2)/(2*scale^2) - log(scale) - log(sqrt(2*pi))."""
    var = scale ** 2
    return -((value - loc) ** 2) / (2 * var) - scale.log() - LOG_SQRT_2PI


def gamma_log_prob(value, concentration, rate):
    """torch/distributions/gamma.py:89-98 (rate parameterisation)."""
    return (torch.xlogy(concentration, rate) + torch.xlogy(concentration - 1, value)
            - rate * value - torch.lgamma(concentration))


def beta_log_prob(value, con1, con0):
    """torch/distributions/beta.py:87-91 -> dirichlet.py:90-97 on stack([v, 1-v])."""
    heads_tails = torch.stack([value, 1.0 - value], -1)
    conc = torch.stack([con1, con0], -1)
    conc, heads_tails = torch.broadcast_tensors(conc, heads_tails)
    return (torch.xlogy(conc - 1.0, heads_tails).sum(-1)
            + torch.lgamma(conc.sum(-1)) - torch.lgamma(conc).sum(-1))


def onehot_categorical_log_prob(value_onehot, probs):
    """one_hot_categorical.py:111-115 -> categorical.py:56-80,151-157 -> utils.py:101-137:
    log(clamp(p / sum p, eps, 1-eps)) gathered at argmax(value) (quirk Q4)."""
    probs = probs / probs.sum(-1, keepdim=True)
    eps = torch.finfo(probs.dtype).eps
    logits = torch.log(probs.clamp(min=eps, max=1 - eps))
    idx = value_onehot.max(-1)[1].unsqueeze(-1)
    idx, logits = torch.broadcast_tensors(idx, logits)
    return logits.gather(-1, idx[..., :1]).squeeze(-1)

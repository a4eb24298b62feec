"""Random draws (they stay in PyTorch, north_star).  `torch.normal(mean_tensor, std_tensor)` -- what
Normal(loc, scale).sample() calls -- checks `std.min() >= 0` on the host, i.e. one device synchronisation per draw and
an operation that cannot be captured in a CUDA graph.  The same arithmetic (standard normal from the current generator,
times scale, plus loc) without the check:"""
import torch


def normal_sample(loc, scale):
    """Normal(loc, scale).sample(): a draw without an autograd graph (the reference never uses rsample)."""
    loc, scale = loc.detach(), scale.detach()
    return torch.randn_like(loc).mul_(scale).add_(loc)

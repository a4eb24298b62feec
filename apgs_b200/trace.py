"""Minimal stand-in for the `probtorch.Trace` bookkeeping the reference's encoder / decoder
contract relies on (apgs/gmm/models.py:25,63,103,140,153,168): an ordered
dict of nodes exposing `.value`, `.dist` and a lazily evaluated, un-summed `.log_prob`.
It carries no arithmetic of its own; the sweep objectives in this package bypass `.log_prob`
and feed the distribution parameters to the fused CUDA kernels instead.
"""
import torch


class Node(object):
    __slots__ = ("dist", "value", "_log_prob")

    def __init__(self, dist, value):
        self.dist = dist
        self.value = value
        self._log_prob = None

    @property
    def log_prob(self):
        if self._log_prob is None:
            self._log_prob = self.dist.log_prob(self.value)
        return self._log_prob


class Trace(dict):
    def variable(self, Dist, *args, name=None, value=None, **kwargs):
        dist = Dist(*args, **kwargs)
        if isinstance(value, Node):
            value = value.value
        if value is None:
            value = dist.rsample() if dist.has_rsample else dist.sample()
        self[name] = Node(dist, value)
        return value

    def normal(self, *args, **kwargs):
        return self.variable(torch.distributions.Normal, *args, **kwargs)

    def gamma(self, *args, **kwargs):
        return self.variable(torch.distributions.Gamma, *args, **kwargs)

    def beta(self, *args, **kwargs):
        return self.variable(torch.distributions.Beta, *args, **kwargs)

"""Import shim for running the UNMODIFIED reference (/root/reference) on CPU.

Used only by tests/golden/make_golden.py and by tests that are skipped when
/root/reference is absent (it does not exist on the GPU box).  Two arithmetic-free stubs
are injected:
  * torch._six      -- removed from modern torch; the reference only reads `inf`
                       (apgs/gmm/kls_gmm.py:171).
  * probtorch.Trace -- un-vendored bookkeeping dependency: an ordered dict of nodes
                       {dist, value, log_prob = dist.log_prob(value)}.
"""
import math
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("APGS_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "apgs"))


class _Node:
    def __init__(self, dist, value):
        self.dist = dist
        self.value = value
        self.log_prob = dist.log_prob(value)


class _Trace(dict):
    def variable(self, Dist, *args, name=None, value=None, **kwargs):
        dist = Dist(*args, **kwargs)
        if isinstance(value, _Node):
            value = value.value
        if value is None:
            value = dist.rsample() if dist.has_rsample else dist.sample()
        self[name] = _Node(dist, value)
        return value

    def normal(self, *args, **kwargs):
        return self.variable(torch.distributions.Normal, *args, **kwargs)

    def gamma(self, *args, **kwargs):
        return self.variable(torch.distributions.Gamma, *args, **kwargs)

    def beta(self, *args, **kwargs):
        return self.variable(torch.distributions.Beta, *args, **kwargs)


def install():
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if "torch._six" not in sys.modules:
        six = types.ModuleType("torch._six")
        six.inf = math.inf
        sys.modules["torch._six"] = six
    if "probtorch" not in sys.modules:
        pt = types.ModuleType("probtorch")
        pt.Trace = _Trace
        sys.modules["probtorch"] = pt
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

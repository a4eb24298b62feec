"""The UNMODIFIED reference (hao-w/apgs) as the CPU arm of bench.py and as a second checker for the GPU tests.

TEST INFRASTRUCTURE ONLY (checker and CPU benchmark arm) -- see oracle/__init__.py.  Nothing under apgs_b200/ imports this.

The reference is a pure-Python package without setup.py / pyproject.toml, so the documented
`pip install --target ... /root/reference` fails ("Neither 'setup.py' nor 'pyproject.toml' found", recorded in
DESIGN.md).  `install()` therefore does what that install would do: it places the package's own files (`apgs/**/*.py`
and the one data asset the bmnist one-shot reads, `apgs/bmnist/mnist_mean.npy`) under `oracle/_ref/` -- a git-ignored
build output like the compiled .so (never committed: the repo history holds no reference source), which travels to
the GPU box with the snapshot.  `__graft_entry__.build()` calls it whenever /root/reference is present.

`load()` puts `oracle/_ref` on sys.path together with the two arithmetic-free import stubs the reference needs under a
modern torch:
  * torch._six      -- removed from torch; apgs/gmm/kls_gmm.py:171 only reads `inf` from it;
  * probtorch.Trace -- un-vendored bookkeeping dependency (ordered dict of nodes {dist, value, log_prob}).
"""
import math
import os
import shutil
import sys
import time
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DST = os.path.join(HERE, "_ref")
REF_SRC = os.environ.get("APGS_REFERENCE", "/root/reference")
_ASSETS = (os.path.join("apgs", "bmnist", "mnist_mean.npy"),)


def install(src=REF_SRC, dst=REF_DST):
    """copy the reference package (python files + mnist_mean.npy) to oracle/_ref/.  Returns the number of files."""
    pkg = os.path.join(src, "apgs")
    if not os.path.isdir(pkg):
        return 0
    n = 0
    for root, _dirs, files in os.walk(pkg):
        rel = os.path.relpath(root, src)
        for f in files:
            relf = os.path.join(rel, f)
            if f.endswith(".py") or relf in _ASSETS:
                os.makedirs(os.path.join(dst, rel), exist_ok=True)
                shutil.copyfile(os.path.join(root, f), os.path.join(dst, relf))
                n += 1
    with open(os.path.join(dst, "INSTALLED_FROM"), "w") as fh:
        fh.write("%s (unmodified files of hao-w/apgs; build output, not tracked)\n" % src)
    return n


def root():
    """directory holding the reference package: oracle/_ref when installed, else the read-only tree (build container)."""
    if os.path.isdir(os.path.join(REF_DST, "apgs")):
        return REF_DST
    if os.path.isdir(os.path.join(REF_SRC, "apgs")):
        return REF_SRC
    return None


def available():
    return root() is not None


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


def load():
    """make `import apgs...` resolve to the reference.  Raises when it is not installed."""
    r = root()
    if r is None:
        raise RuntimeError("reference not installed: run __graft_entry__.build() where /root/reference exists")
    if "torch._six" not in sys.modules:
        six = types.ModuleType("torch._six")
        six.inf = math.inf
        sys.modules["torch._six"] = six
    if "probtorch" not in sys.modules:
        pt = types.ModuleType("probtorch")
        pt.Trace = _Trace
        sys.modules["probtorch"] = pt
    if r not in sys.path:
        sys.path.insert(0, r)
    import warnings
    warnings.filterwarnings("ignore", category=SyntaxWarning)      # the reference's docstrings hold stray backslashes
    return r


# ------------------------------------------------------------------------------------------------ timed workloads
FLAGS = {'loss_required': False, 'ess_required': True, 'mode_required': False, 'density_required': True}


def _time(fn, steps, warmup, sync=None):
    for _ in range(warmup):
        fn()
    if sync:
        sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    if sync:
        sync()
    return (time.perf_counter() - t0) / max(1, steps)


def gmm_step_fn(S, B, sweeps, device="cpu", K=3, D=2, N=60, H=32, seed=0, strategy="systematic", block="decomposed"):
    """closure running the reference's own apg_objective (apgs/gmm/objectives.py:9-53) on synthetic data of the
    simulator's distribution, models from its init_apg_models (apgs/gmm/apg_training.py:70-99), eval flags of
    budget_analysis (apgs/gmm/evaluation.py:127)."""
    load()
    from apgs.gmm.apg_training import init_apg_models
    from apgs.gmm.objectives import apg_objective
    from apgs.resampler import Resampler
    from . import gmm as og
    cuda = device != "cpu"
    dev = torch.device(device)
    torch.manual_seed(seed)
    models = init_apg_models(K, D, H, cuda, dev)
    x = og.simulate(B, N, K, D).to(dev).repeat(S, 1, 1, 1)
    rs = Resampler(strategy, S, cuda, dev)

    def step():
        with torch.no_grad():
            return apg_objective(models, x, FLAGS, sweeps, block, rs)
    return step


def dmm_step_fn(S, B, sweeps, device="cpu", K=4, D=2, N=200, seed=0):
    """the reference's DMM apg_objective (apgs/dmm/objectives.py:9-56), models from apgs/dmm/apg_training.py init."""
    load()
    from apgs.dmm.apg_training import init_apg_models
    from apgs.dmm.objectives import apg_objective
    from apgs.resampler import Resampler
    from . import dmm as od
    cuda = device != "cpu"
    dev = torch.device(device)
    torch.manual_seed(seed)
    models = init_apg_models(K, D, 32, 8, 32, 32, 0.5, cuda, dev)
    x = od.simulate(B, N, K, D).to(dev).repeat(S, 1, 1, 1)
    rs = Resampler("systematic", S, cuda, dev)

    def step():
        with torch.no_grad():
            return apg_objective(models, x, K, FLAGS, sweeps, rs)
    return step


def bmnist_step_fn(S, B, sweeps, device="cpu", K=3, T=10, seed=0):
    """the reference's bmnist apg_objective (apgs/bmnist/objectives.py:11-58)."""
    load()
    import numpy as np
    from apgs.bmnist.apg_training import init_models
    from apgs.bmnist.objectives import apg_objective
    from apgs.bmnist.affine_transformer import Affine_Transformer
    from apgs.resampler import Resampler
    from . import bmnist as ob
    cuda = device != "cpu"
    dev = torch.device(device)
    torch.manual_seed(seed)
    models = init_models(96, 28, 400, 400, 2, 10, cuda, dev, None, None)
    AT = Affine_Transformer(96, 28, cuda, dev)
    mean = torch.from_numpy(np.load(os.path.join(root(), "apgs", "bmnist", "mnist_mean.npy"))).float()
    frames = ob.simulate(B, T, K).to(dev).repeat(S, 1, 1, 1, 1)
    mnist_mean = mean.to(dev).repeat(S, B, K, 1, 1)
    rs = Resampler("systematic", S, cuda, dev)

    def step():
        with torch.no_grad():
            return apg_objective(models, AT, frames, K, FLAGS, sweeps, rs, mnist_mean)
    return step


def resample_step_fn(S, B, d, strategy="systematic", device="cpu", sigma=0.1, seed=0):
    """the reference's resample of ONE latent (S,B,1,d): sample_ancestral_index + resample_4dims (apgs/resampler.py:22-49)."""
    load()
    from apgs.resampler import Resampler
    cuda = device != "cpu"
    dev = torch.device(device)
    torch.manual_seed(seed)
    rs = Resampler(strategy, S, cuda, dev)
    lw = (torch.randn(S, B) * sigma).to(dev)
    var = torch.randn(S, B, 1, d).to(dev)

    def step():
        a = rs.sample_ancestral_index(lw)
        return rs.resample_4dims(var, a)
    return step


def throughput(step, units_per_step, steps, warmup, threads=None, sync=None):
    """(units/s, seconds per step) of `step` with torch's intra-op pool set to `threads` (default: all host cores)."""
    torch.set_num_threads(threads or (os.cpu_count() or 1))
    sec = _time(step, steps, warmup, sync)
    return units_per_step / sec, sec

"""Training step (objective + backward) vs evaluation step at the GMM shapes: the kernel-backed graphs (train_on_kernels) against
the PyTorch-encoder graphs.  python benchmarks/train_step.py [S B sweeps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apgs_b200.gmm import models as M, objectives as O  # noqa: E402
from apgs_b200.gmm.sim import simulate  # noqa: E402
from apgs_b200.resampler import Resampler  # noqa: E402

DEV = "cuda:0"
S, B, sweeps = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (1000, 1024, 10)))
EVAL = {'loss_required': False, 'ess_required': True, 'mode_required': False, 'density_required': True}
TRAIN = dict(EVAL, loss_required=True)


def timed(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


torch.manual_seed(0)
models, opt = M.init_apg_models(3, 2, 32, True, DEV, lr=1e-4)
x = simulate(B).to(DEV)
rs = Resampler('systematic', S, True, DEV)


def train_step():
    opt.zero_grad()
    tr = O.apg_objective(models, x, TRAIN, sweeps, 'decomposed', rs)
    tr['loss'].sum().backward()
    opt.step()


def eval_step():
    with torch.no_grad():
        O.apg_objective(models, x, EVAL, sweeps, 'decomposed', rs)


row = {"S": S, "B": B, "sweeps": sweeps}
row["eval_ms"] = timed(eval_step)
O.train_on_kernels = True
row["train_kernels_ms"] = timed(train_step)
row["peak_GB_kernels"] = torch.cuda.max_memory_allocated() / 1e9
try:
    torch.cuda.reset_peak_memory_stats()
    O.train_on_kernels = False
    row["train_torch_ms"] = timed(train_step, iters=2)
    row["peak_GB_torch"] = torch.cuda.max_memory_allocated() / 1e9
except Exception as e:                                              # noqa: BLE001
    row["train_torch_error"] = repr(e)[:200]
row["train_over_eval"] = row["train_kernels_ms"] / row["eval_ms"]
print(json.dumps(row))

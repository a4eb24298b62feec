"""per-kernel CUDA time of one training step (torch profiler): python benchmarks/_train_prof.py [S B sweeps]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apgs_b200.gmm import models as M, objectives as O  # noqa: E402
from apgs_b200.gmm.sim import simulate  # noqa: E402
from apgs_b200.resampler import Resampler  # noqa: E402

DEV = "cuda:0"
S, B, sweeps = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (1000, 256, 10)))
TRAIN = {'loss_required': True, 'ess_required': True, 'mode_required': False, 'density_required': True}
torch.manual_seed(0)
models, opt = M.init_apg_models(3, 2, 32, True, DEV, lr=1e-4)
x = simulate(B).to(DEV)
rs = Resampler('systematic', S, True, DEV)


def train_step():
    opt.zero_grad()
    tr = O.apg_objective(models, x, TRAIN, sweeps, 'decomposed', rs)
    tr['loss'].sum().backward()
    opt.step()


train_step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    train_step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))

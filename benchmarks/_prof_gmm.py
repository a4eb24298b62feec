import os, sys, torch
sys.path.insert(0, os.getcwd())
from apgs_b200.gmm import apg_objective, init_apg_models
from apgs_b200.gmm.sim import simulate
from apgs_b200.resampler import Resampler
dev = torch.device("cuda", 0)
S, B, sweeps = 1000, 1024, 4
torch.manual_seed(0)
models = init_apg_models(3, 2, 32, True, dev)
x = simulate(B).to(dev)
flags = {'loss_required': False, 'ess_required': True, 'mode_required': False, 'density_required': True}
rs = Resampler('systematic', S, True, dev)
with torch.no_grad():
    for _ in range(2):
        apg_objective(models, x, flags, sweeps, 'decomposed', rs)
torch.cuda.synchronize()

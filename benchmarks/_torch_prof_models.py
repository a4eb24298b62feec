"""torch profiler breakdown of one DMM (C3) / bmnist (C4) objective: what is left outside our kernels"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from torch.profiler import profile, ProfilerActivity
from apgs_b200.resampler import Resampler
FLAGS = {'loss_required': False, 'ess_required': True, 'mode_required': False, 'density_required': True}
DEV = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "dmm"
torch.manual_seed(0)
if which == "dmm":
    from apgs_b200.dmm import apg_objective, init_apg_models
    from oracle import dmm as od
    K, D, N, S, B, sweeps = 4, 2, 200, 100, 256, 7
    models = init_apg_models(K, D, 32, 8, 32, 32, 0.5, True, DEV)
    xd = od.simulate(B, N, K, D).to(DEV)
    rs = Resampler('systematic', S, True, DEV)
    fn = lambda: apg_objective(models, xd, K, FLAGS, sweeps, rs)
else:
    from apgs_b200.bmnist import apg_objective, init_models, Affine_Transformer
    from oracle import bmnist as ob
    K, FP, DP, S, B, T, sweeps = 3, 96, 28, 10, 64, 10, 5
    models = init_models(FP, DP, 400, 400, 2, 10, True, DEV)
    frames = ob.simulate(B, T, K)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, DP), torch.linspace(-1, 1, DP), indexing='ij')
    tmpl = torch.exp(-(xx ** 2 + yy ** 2) * 3).clamp(0, 1)
    AT = Affine_Transformer(FP, DP, True, DEV)
    rs = Resampler('systematic', S, True, DEV)
    fd = frames.to(DEV).unsqueeze(0).expand(S, -1, -1, -1, -1)
    mm = tmpl.to(DEV).repeat(S, B, K, 1, 1)
    fn = lambda: apg_objective(models, AT, fd, K, FLAGS, sweeps, rs, mm)
with torch.no_grad():
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print("step ms", e0.elapsed_time(e1))
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))

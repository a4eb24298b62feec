"""minimal driver for ncu captures of the bmnist template-matching kernel at the C4 shape (S*B = 640 images)"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from apgs_b200.bmnist import kernels
torch.manual_seed(0)
S, B = 10, 64
f = (torch.rand(S, B, 96, 96, device="cuda") > 0.8).float()
t = torch.rand(S, B, 28, 28, device="cuda")
for _ in range(4):
    out = kernels.template_match(f, t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = kernels.template_match(f, t)
e1.record(); torch.cuda.synchronize()
print("template_match ms", e0.elapsed_time(e1) / 10)

import torch, sys, os
sys.path.insert(0, os.getcwd())
from apgs_b200 import ops
dev='cuda:0'
S,B=1000,1024
torch.manual_seed(0)
lw=torch.randn(S,B,device=dev)*2
u=torch.rand(B,device=dev)
tau=torch.randn(S,B,3,2,device=dev); mu=torch.randn(S,B,3,2,device=dev)
z=torch.nn.functional.one_hot(torch.randint(0,3,(S,B,60),device=dev),3).float()
for _ in range(3):
    outs,_,_=ops.resample([tau,mu,z], log_w=lw, strategy='systematic', uniforms=u)
torch.cuda.synchronize()

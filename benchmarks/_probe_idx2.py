import torch, sys, os
sys.path.insert(0, os.getcwd())
from apgs_b200 import ops, _lib
dev='cuda:0'
S,B=1000,1024
lw=torch.randn(S,B,device=dev)*2
u=torch.rand(B,device=dev)
w=torch.softmax(lw,0)
for _ in range(3):
    ops.weight_stats(lw, want_w=False)
    ops.sample_ancestors(log_w=lw, uniforms=u, index_dtype=torch.int32)
    ops.sample_ancestors(log_w=lw, uniforms=u, index_dtype=torch.int32, cdf_mode=_lib.CDF_BLOCKED)
    ops.sample_ancestors(w_norm=w, uniforms=u, index_dtype=torch.int32)
    um=torch.rand(S,B,dtype=torch.float64,device=dev)
    ops.sample_ancestors(log_w=lw, strategy='multinomial', uniforms=um, index_dtype=torch.int32)
torch.cuda.synchronize()

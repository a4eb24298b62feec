import torch, sys, os, time
sys.path.insert(0, os.getcwd())
from apgs_b200 import ops, _lib
from benchmarks.resample_microbench import timed
dev='cuda:0'
for (S,B) in [(1000,1024),(100,256),(1000,64)]:
    lw=torch.randn(S,B,device=dev)*2
    u=torch.rand(B,device=dev)
    w=torch.softmax(lw,0)
    t_stats=timed(lambda: ops.weight_stats(lw, want_w=False), 30)
    t_exact=timed(lambda: ops.sample_ancestors(log_w=lw, uniforms=u, index_dtype=torch.int32), 30)
    t_blk=timed(lambda: ops.sample_ancestors(log_w=lw, uniforms=u, index_dtype=torch.int32, cdf_mode=_lib.CDF_BLOCKED), 30)
    t_wn=timed(lambda: ops.sample_ancestors(w_norm=w, uniforms=u, index_dtype=torch.int32), 30)
    t_empty=timed(lambda: torch.empty(S,B,device=dev), 30)
    # host overhead: time N back-to-back calls wall-clock
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(200): ops.sample_ancestors(log_w=lw, uniforms=u, index_dtype=torch.int32)
    torch.cuda.synchronize(); t1=time.perf_counter()
    print(S,B,'stats %.4f exact %.4f blocked %.4f from_w %.4f empty %.4f ms ; back-to-back %.4f ms/call'%(t_stats,t_exact,t_blk,t_wn,t_empty,(t1-t0)/200*1e3))

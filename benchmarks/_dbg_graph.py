import os, sys, torch
sys.path.insert(0, os.getcwd())
from apgs_b200 import ops
from apgs_b200.gmm import kernels, init_apg_models
from apgs_b200.resampler import Resampler
from torch.distributions.gamma import Gamma
from torch.distributions.normal import Normal
dev=torch.device('cuda',0)
S,B,N,K,D=10,20,60,3,2
torch.manual_seed(0)
models=init_apg_models(K,D,32,True,dev)
enc_rws_eta, enc_apg_z, enc_apg_eta, gen = models
x=torch.randn(B,N,D,device=dev)
z=torch.nn.functional.one_hot(torch.randint(0,K,(S,B,N),device=dev),K).float()
lw=torch.randn(S,B,device=dev)
tau=torch.rand(S,B,K,D,device=dev)+.5; mu=torch.randn(S,B,K,D,device=dev)
rs=Resampler('systematic',S,True,dev)
post=enc_apg_eta.params(x,z,gen.prior_ng)
def try_capture(name, fn):
    s=torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g=torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print('OK  ', name)
    except Exception as e:
        print('FAIL', name, repr(e)[:120])
        torch.cuda.synchronize()
with torch.no_grad():
    try_capture('weight_stats', lambda: ops.weight_stats(lw))
    try_capture('rand', lambda: torch.rand(B,1,device=dev))
    try_capture('resample', lambda: rs.resample([tau,mu,z], lw))
    try_capture('eta_post_fused', lambda: enc_apg_eta.params(x,z,gen.prior_ng))
    try_capture('rws_params', lambda: enc_rws_eta.params(x.unsqueeze(0).expand(S,-1,-1,-1),gen.prior_ng))
    try_capture('gamma', lambda: Gamma(post[0],post[1],validate_args=False).sample())
    try_capture('normal', lambda: Normal(post[2],post[3],validate_args=False).sample())
    try_capture('logw_eta', lambda: kernels.logw_eta(x,z,post,gen,tau,mu,tau,mu))
    try_capture('z_fused', lambda: kernels.z_step_fused(x,tau,mu,enc_apg_z.mlp(),gen,z_old=z))
    try_capture('z_fused_sample', lambda: enc_apg_z.sample_fused(x,tau,mu,gen))
    try_capture('logw_oneshot', lambda: kernels.logw_oneshot(x,post,gen,tau,mu,z,z))

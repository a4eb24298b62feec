import torch
dev='cuda:0'
def timed(fn, n=20):
    for _ in range(3): fn()
    ts=[]
    for _ in range(n):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts)//2]
x=torch.empty(788_000_000//4, device=dev); y=torch.empty_like(x)
t=timed(lambda: x.fill_(1.0)); print('fill 788MB  %.3f ms  %.0f GB/s'%(t, 788/t))
t=timed(lambda: x.zero_()); print('zero 788MB  %.3f ms  %.0f GB/s'%(t, 788/t))
t=timed(lambda: y.copy_(x)); print('copy 788MB  %.3f ms  %.0f GB/s (r+w)'%(t, 2*788/t))
t=timed(lambda: x.sum()); print('sum  788MB  %.3f ms  %.0f GB/s (read)'%(t, 788/t))

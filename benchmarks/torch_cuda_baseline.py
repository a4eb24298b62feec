"""The reference's algorithm through STOCK PyTorch CUDA ops ("also time the reference on the same
GPU through stock PyTorch CUDA ops as a second, stronger baseline").  /root/reference does not travel to the GPU box,
so this runs the oracle's restatement of it (bit-identical to the reference on CPU) with every tensor on cuda:0 and a
resampler written with the reference's own ops (softmax, cumsum, searchsorted, expanded-index gather:
apgs/resampler.py:22-54).  A benchmark baseline only -- nothing in apgs_b200/ uses it.

    python benchmarks/torch_cuda_baseline.py [--S 1000] [--B 128] [--sweeps 10] [--out gpurun_out/torch_cuda.json]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gmm as og  # noqa: E402

DEV = torch.device("cuda", 0)
FLAGS = {'loss_required': False, 'ess_required': True, 'mode_required': False, 'density_required': True}


class TorchCudaResampler:
    """apgs/resampler.py:22-54 with the reference's ops, on the GPU."""

    def __init__(self, strategy, S):
        self.strategy, self.S = strategy, S
        self.spacing = torch.arange(S, device=DEV).float()

    def ancestors(self, log_w):
        S, B = log_w.shape
        w = F.softmax(log_w, 0).t()
        if self.strategy == 'systematic':
            positions = (torch.rand(B, 1, device=DEV) + self.spacing) / S
            cdf = torch.cumsum(w, 1)
            cdf = cdf / cdf.max(1, keepdim=True).values
            return torch.searchsorted(cdf, positions.contiguous()).t().clamp_(max=S - 1)
        return torch.multinomial(w, S, replacement=True).t()

    def gather(self, var, a):
        idx = a.reshape(a.shape + (1,) * (var.dim() - 2)).repeat(1, 1, *var.shape[2:])      # materialised, as the reference
        return torch.gather(var, 0, idx)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--S", type=int, default=1000)
    ap.add_argument("--B", type=int, default=128)
    ap.add_argument("--sweeps", type=int, default=10)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    torch.manual_seed(0)
    P = {k: {n: v.to(DEV) for n, v in d.items()} for k, d in og.make_params().items()}
    prior = {k: v.to(DEV) for k, v in og.make_prior().items()}
    x = og.simulate(a.B).to(DEV).repeat(a.S, 1, 1, 1)            # S-replicated, as the reference passes it
    og._Resampler = TorchCudaResampler
    run = lambda: og.apg_objective(P, prior, x, FLAGS, a.sweeps, 'decomposed', 'systematic', mirror_cost=True)
    with torch.no_grad():
        run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) / 1e3)
    t = sorted(ts)[len(ts) // 2]
    row = {"impl": "oracle port of the reference on stock PyTorch CUDA ops", "S": a.S, "B": a.B, "sweeps": a.sweeps,
           "seconds": t, "particle_sweeps_per_s": a.S * a.B * a.sweeps / t,
           "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}
    print(json.dumps(row))
    if a.out:
        json.dump(row, open(a.out, "w"))


if __name__ == "__main__":
    main()

"""CUDA-graph replay of a whole APG objective (B200-first plumbing: streams and graphs instead of a tracing compiler).

At the reference's own training shapes (S=10, B=20) one objective is ~250 launches of microsecond kernels and is bound
by launch latency; captured once and replayed it costs a single graph launch.  Everything on the path is capturable:
the C-ABI kernels launch on torch's current (capturing) stream, allocate nothing, and the random draws use PyTorch's
graph-safe Philox generator (fresh numbers on every replay).

    g = GraphedObjective(lambda x: apg_objective(models, x, flags, sweeps, 'decomposed', resampler), x_example)
    trace = g(x)            # copies x into the static input, replays, returns the (static) output tensors
"""
import torch


def _map(obj, fn):
    if torch.is_tensor(obj):
        return fn(obj)
    if isinstance(obj, dict):
        return {k: _map(v, fn) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_map(v, fn) for v in obj)
    return obj


class GraphedObjective(object):
    def __init__(self, fn, *example_inputs, warmup=3):
        if not all(torch.is_tensor(t) and t.is_cuda for t in example_inputs):
            raise ValueError("example inputs must be CUDA tensors")
        self.static_in = [t.detach().clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):                       # also sets every kernel's shared-memory attribute before capture
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs, clone=False):
        for dst, src in zip(self.static_in, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return _map(self.static_out, lambda t: t.clone()) if clone else self.static_out

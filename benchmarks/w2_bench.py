"""W2 (per-point softmax over S, apgs/gmm/objectives.py:216-218) at the C2 / C3 shapes: ms and achieved HBM GB/s
(algorithmic bytes = one read + one write of the (S,B,N) tensor)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apgs_b200 import ops  # noqa: E402

PEAK = 6552.0
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


for S, B, N in ((1000, 1024, 60), (100, 256, 200)):
    x = torch.randn(S, B, N, device="cuda") * 3
    t = timed(lambda: ops.local_weight_stats(x, want_w=True, want_sum=False))
    t_torch = timed(lambda: torch.softmax(x, 0))
    t_sum = timed(lambda: ops.local_weight_stats(x, want_w=False, want_sum=True))
    nbytes = 2 * x.numel() * 4
    print(json.dumps({"S": S, "B": B, "N": N, "softmax_ms": round(t, 4), "softmax_GBs": round(nbytes / t / 1e6, 1),
                      "softmax_frac_hbm": round(nbytes / t / 1e6 / PEAK, 3), "torch_softmax_ms": round(t_torch, 4),
                      "sum_ms": round(t_sum, 4), "sum_GBs": round(x.numel() * 4 / t_sum / 1e6, 1)}))

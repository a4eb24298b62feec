"""Resampler microbench (BASELINE.json configs[4] / SURVEY.md C5): systematic vs multinomial over
P = S*B particles x latent dim d, kernel times by CUDA events, achieved algorithmic GB/s.

    python benchmarks/resample_microbench.py [--quick] [--gmm]

Algorithmic bytes per particle: 4 (log_w) + 2 * inner_bytes per latent.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apgs_b200 import ops  # noqa: E402
from apgs_b200.resampler import Resampler  # noqa: E402

PEAK = 6552.0
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, iters=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def run_case(S, B, inners, strategy, sigma, flush, iters=20):
    dev = "cuda:0"
    lw = torch.randn(S, B, device=dev) * sigma
    lat = [torch.randn(S, B, max(1, n // 4), device=dev) for n in inners]
    r = Resampler(strategy, S, True, dev)
    u = r.draw_uniforms(S, B, dev)
    lay = "bs" if strategy == "multinomial" else None           # draw_uniforms: (B,S), the reference's generator order
    anc = ops.sample_ancestors(log_w=lw, strategy=strategy, uniforms=u, index_dtype=torch.int32, uniform_layout=lay)
    t_idx = timed(lambda: ops.sample_ancestors(log_w=lw, strategy=strategy, uniforms=u, index_dtype=torch.int32,
                                               uniform_layout=lay), iters, flush=flush)
    t_gat = timed(lambda: ops.gather_ancestors(lat, anc), iters, flush=flush)
    t_all = timed(lambda: ops.resample(lat, log_w=lw, strategy=strategy, uniforms=u, uniform_layout=lay), iters, flush=flush)
    P = S * B
    bytes_g = P * (4 + 2 * sum(inners))
    bytes_all = P * (4 + 2 * sum(inners))
    return {"S": S, "B": B, "inner_bytes": inners, "strategy": strategy, "sigma": sigma,
            "ms_index": round(t_idx, 4), "ms_gather": round(t_gat, 4), "ms_resample": round(t_all, 4),
            "gather_GBs": round(bytes_g / t_gat / 1e6, 1), "resample_GBs": round(bytes_all / t_all / 1e6, 1),
            "frac_of_measured_peak": round(bytes_all / t_all / 1e6 / PEAK, 3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--gmm", action="store_true", help="only the GMM S=1000 B=1024 latents (tau, mu, z)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda:0")   # > 126 MB L2
    rows = []
    if a.gmm or True:
        for strat in ("systematic", "multinomial"):
            for sigma in (0.1, 2.0, 10.0):            # ESS ~ S, ~ S/50, ~ 1: distinct ancestors read from HBM vary
                rows.append(run_case(1000, 1024, [24, 24, 720], strat, sigma, flush))
                print(json.dumps(rows[-1]), flush=True)
        rows.append(run_case(100, 256, [32, 3200, 800], "systematic", 2.0, flush))   # DMM C3
        print(json.dumps(rows[-1]), flush=True)
    if not a.gmm:
        Ps = [10**4, 10**6] if a.quick else [10**4, 10**5, 10**6, 10**7, 10**8]
        for P in Ps:
            for S in (100, 1000, 10000, 100000):
                B = P // S
                if B < 1:
                    continue
                for d in ((2, 32, 256) if a.quick else (2, 8, 32, 128, 256)):
                    if P * d * 4 * 2 > 60e9:
                        continue
                    for strat in ("systematic", "multinomial"):
                        rows.append(run_case(S, B, [4 * d], strat, 2.0, flush, iters=10))
                        print(json.dumps(rows[-1]), flush=True)
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()

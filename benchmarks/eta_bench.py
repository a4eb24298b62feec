"""The eta step's kernels at the C2 shape (label mode): statistics + posterior (k_gmm_nss_im), fused draw + weights:
ms per launch.  APGS_LIB selects another build of the library."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apgs_b200 import _lib  # noqa: E402
if os.environ.get("APGS_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["APGS_LIB"])
from apgs_b200.gmm import kernels, models as M  # noqa: E402
from apgs_b200.gmm.sim import simulate  # noqa: E402

DEV = "cuda:0"
S, B, N, K, D = 1000, 1024, 60, 3, 2


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


torch.manual_seed(0)
(_, enc_z, enc_eta, gen) = M.init_apg_models(K, D, 32, True, DEV)
x = simulate(B).to(DEV)
tau, mu = torch.rand(S, B, K, D, device=DEV) * 3 + 0.2, torch.randn(S, B, K, D, device=DEV) * 2
lab = torch.randint(0, K, (S, B, 64), dtype=torch.uint8, device=DEV)
heads = enc_eta.head_weights()
out = kernels.eta_posterior_fused(x, lab, heads, gen, center=mu)
post, stats = out[:4], out[4]
seed = kernels.philox_seed(DEV)
row = {"lib": os.path.basename(_lib.LIB_PATH),
       "eta_posterior_stats_ms": timed(lambda: kernels.eta_posterior_fused(x, lab, heads, gen, center=mu)),
       "draw_logw_eta_ms": timed(lambda: kernels.draw_logw_eta_stats_rng(stats, mu, post, gen, seed, tau, mu))}
print(json.dumps(row))

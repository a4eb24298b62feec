"""k_gmm_z_logq_bwd at the C2 shape: ms per launch (+ the forward launch that feeds it, and the label-table pair)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apgs_b200.gmm import kernels, models as M, train  # noqa: E402
from apgs_b200.gmm.sim import simulate  # noqa: E402

DEV = "cuda:0"
S, B, N, K, D = 1000, 1024, 60, 3, 2


def timed(fn, iters=10, warm=2):
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
z, q, w = kernels.z_step_fused(x, tau, mu, enc_z.mlp(), gen, z_old=lab, want_q=True, local=True)
G = torch.randn(S, B, N, device=DEV)
row = {"z_fwd_train_ms": timed(lambda: kernels.z_step_fused(x, tau, mu, enc_z.mlp(), gen, z_old=lab, want_q=True, local=True)),
       "z_bwd_ms": timed(lambda: train.z_logq_backward(x, tau, mu, enc_z.mlp(), q, z, G))}
tab = torch.randn(B, N, K, 15, device=DEV)
g = torch.randn(S, B, 15, device=DEV)
row["lts_fwd_ms"] = timed(lambda: train.label_table_sum(tab, lab))
row["lts_bwd_ms"] = timed(lambda: train.label_table_sum_backward(g, lab, N, K))
terms = S * B * N * K * 32
row["z_bwd_Tterms_per_s"] = terms / row["z_bwd_ms"] / 1e9
print(json.dumps(row))

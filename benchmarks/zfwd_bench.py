"""The lean label-mode z step (in-kernel uniforms) at the C2 shape: ms per launch.  APGS_LIB selects another build of
the library (measurement-only variants, e.g. -DAPGS_ZF_EXPT=1)."""
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
S, B, N, K, D = 1000, int(os.environ.get("ZB", 1024)), 60, 3, 2


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
seed = kernels.philox_seed(DEV)
u = torch.rand(S, B, N, device=DEV)
tab = kernels.z_tables(x, enc_z.mlp(), gen) if hasattr(kernels, "z_tables") and not os.environ.get("NO_TAB") else None
row = {"lib": os.path.basename(_lib.LIB_PATH),
       "z_lean_rng_ms": timed(lambda: kernels.z_step_fused(x, tau, mu, enc_z.mlp(), gen, z_old=lab, seed=seed, want_onehot=False)),
       "z_lean_u_ms": timed(lambda: kernels.z_step_fused(x, tau, mu, enc_z.mlp(), gen, z_old=lab, uniforms=u, want_onehot=False))}
if tab is not None:
    row["z_lean_rng_tab_ms"] = timed(lambda: kernels.z_step_fused(x, tau, mu, enc_z.mlp(), gen, z_old=lab, seed=seed, want_onehot=False, tables=tab))
    row["z_tables_ms"] = timed(lambda: kernels.z_tables(x, enc_z.mlp(), gen))
print(json.dumps(row))

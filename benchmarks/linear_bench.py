"""Enc_coor's Linear(4761 -> 400) + ReLU at the C4 shape (M = S*B = 640): the 3xTF32 tensor-core kernel against torch's fp32
linear (cuBLAS SIMT sgemm): ms per call and the error against an fp64 product."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apgs_b200 import _lib  # noqa: E402
if os.environ.get("APGS_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["APGS_LIB"])
from apgs_b200.bmnist import kernels  # noqa: E402

DEV = "cuda:0"
M, K, N = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (640, 4761, 400)))


def timed(fn, iters=30, warm=5):
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


def graph_timed(fn, calls=20, reps=10):
    """ms per call on the device: `calls` calls captured into one CUDA graph (no host launch overhead in the timed region)"""
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(st)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(calls):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) / calls)
    return sorted(ts)[len(ts) // 2]


torch.manual_seed(0)
x = torch.softmax(torch.randn(M, K, device=DEV) * 3, -1)
w = torch.randn(N, K, device=DEV) / K ** 0.5
b = torch.randn(N, device=DEV)
ref = (x.double() @ w.double().t() + b.double()).clamp_min(0)
mag = x.double().abs() @ w.double().abs().t() + b.double().abs()
ours = kernels.linear_tf32x3(x, w, b, relu=True)
th = torch.relu(torch.nn.functional.linear(x, w, b))
tc5 = kernels.linear_tc5(x, w, b, relu=True)
torch.cuda.synchronize()
row = {"M": M, "K": K, "N": N, "tc5_ms": timed(lambda: kernels.linear_tc5(x, w, b, relu=True)),
       "tc5_err": float(((tc5.double() - ref).abs() / mag).max()),
       "tf32x3_ms": timed(lambda: kernels.linear_tf32x3(x, w, b, relu=True)),
       "torch_fp32_ms": timed(lambda: torch.relu(torch.nn.functional.linear(x, w, b))),
       "tf32x3_err": float(((ours.double() - ref).abs() / mag).max()), "torch_fp32_err": float(((th.double() - ref).abs() / mag).max())}
xp, wp = kernels.pack_rows(x, 0), kernels.pack_rows(w, 1)
pk = kernels.linear_tc5_packed(xp, wp, b, relu=True)
row["tc5_packed_ms"] = timed(lambda: kernels.linear_tc5_packed(xp, wp, b, relu=True))
row["graph_ms"] = {"tc5": graph_timed(lambda: kernels.linear_tc5(x, w, b, relu=True)),
                   "tc5_packed": graph_timed(lambda: kernels.linear_tc5_packed(xp, wp, b, relu=True)),
                   "torch_fp32": graph_timed(lambda: torch.relu(torch.nn.functional.linear(x, w, b)))}
row["tc5_packed_equals_tc5"] = bool(torch.equal(pk, tc5))
row["pack_w_ms"] = timed(lambda: kernels.pack_rows(w, 1))
row["pack_x_ms"] = timed(lambda: kernels.pack_rows(x, 0))
row["tc5pk_stages"] = int(os.environ.get("APGS_TC5PK_STAGES", "2"))
# Enc_coor's two heads on the hidden layer (400 -> 200 -> 2, tanh / exp): one launch against the nn.Sequential modules
import torch.nn as nn  # noqa: E402
hm = nn.Sequential(nn.Linear(N, N // 2), nn.ReLU(), nn.Linear(N // 2, 2), nn.Tanh()).to(DEV)
hs = nn.Sequential(nn.Linear(N, N // 2), nn.ReLU(), nn.Linear(N // 2, 2)).to(DEV)
w1t, b1 = kernels.coor_heads_weights(hm[0].weight, hm[0].bias, hs[0].weight, hs[0].bias)
with torch.no_grad():
    row["graph_ms"]["heads_fused"] = graph_timed(lambda: kernels.coor_heads(tc5, w1t, b1, hm[2].weight, hm[2].bias, hs[2].weight, hs[2].bias))
    row["graph_ms"]["heads_torch"] = graph_timed(lambda: (hm(tc5), hs(tc5).exp()))
row["TFLOPs_effective"] = 2.0 * M * K * N / row["tf32x3_ms"] / 1e9
print(json.dumps(row))

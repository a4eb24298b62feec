"""Throughput of the three APG tasks at the BASELINE.json shapes on one GPU, next to the CPU oracle port
(bit-identical to the reference on CPU) on a bounded sample.  particle-sweeps/s = S * B * sweeps / time(apg_objective).

    python benchmarks/bench_models.py [--out gpurun_out/models.json] [--no-cpu]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/bench_models.py --no-cpu
        weak scaling along the instance axis: every rank runs the SAME per-GPU shape on its own instances (the path
        shards along B with no data-path collective); a row then reports the slowest rank's time and N times the work.
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
FLAGS = {'loss_required': False, 'ess_required': True, 'mode_required': False, 'density_required': True}
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
DEV = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))


def gpu_time(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) / 1e3)
    return sorted(ts)[len(ts) // 2]


def cpu_time(fn, iters=1):
    torch.set_num_threads(os.cpu_count() or 1)
    fn()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t0) / iters


def gmm(S, B, sweeps, cpu_B, no_cpu):
    from apgs_b200.gmm import apg_objective, init_apg_models
    from apgs_b200.resampler import Resampler
    from oracle import gmm as og
    torch.manual_seed(0)
    models = init_apg_models(3, 2, 32, True, DEV)
    x = og.simulate(B)
    rs = Resampler('systematic', S, True, DEV)
    xd = x.to(DEV)
    with torch.no_grad():
        t = gpu_time(lambda: apg_objective(models, xd, FLAGS, sweeps, 'decomposed', rs))
    row = {"task": "gmm", "S": S, "B": B, "sweeps": sweeps, "gpu_s": t, "gpu_psw_s": S * B * sweeps / t}
    try:                                            # the same objective captured once into a CUDA graph
        from apgs_b200.graph import GraphedObjective
        g = GraphedObjective(lambda xx: apg_objective(models, xx, FLAGS, sweeps, 'decomposed', rs), xd)
        tg = gpu_time(lambda: g(xd))
        row.update({"graph_s": tg, "graph_psw_s": S * B * sweeps / tg})
    except Exception as e:                          # noqa: BLE001
        row["graph_error"] = repr(e)[:200]
    if not no_cpu:
        P, prior = og.make_params(), og.make_prior()
        xc = x[:cpu_B].repeat(S, 1, 1, 1)
        with torch.no_grad():
            tc = cpu_time(lambda: og.apg_objective(P, prior, xc, FLAGS, sweeps, 'decomposed', 'systematic', mirror_cost=True))
        row.update({"cpu_B": cpu_B, "cpu_s": tc, "cpu_psw_s": S * cpu_B * sweeps / tc, "cpu_threads": torch.get_num_threads()})
    return row


def dmm(S, B, sweeps, cpu_B, no_cpu):
    from apgs_b200.dmm import apg_objective, init_apg_models
    from apgs_b200.resampler import Resampler
    from oracle import dmm as od
    K, D, N = 4, 2, 200
    torch.manual_seed(0)
    models = init_apg_models(K, D, 32, 8, 32, 32, 0.5, True, DEV)
    x = od.simulate(B, N, K, D)
    rs = Resampler('systematic', S, True, DEV)
    xd = x.to(DEV)
    with torch.no_grad():
        t = gpu_time(lambda: apg_objective(models, xd, K, FLAGS, sweeps, rs))
    row = {"task": "dmm", "S": S, "B": B, "sweeps": sweeps, "gpu_s": t, "gpu_psw_s": S * B * sweeps / t}
    if not no_cpu:
        P, prior = od.make_params(K, D), od.make_prior(K, D)
        xc = x[:cpu_B].repeat(S, 1, 1, 1)
        with torch.no_grad():
            tc = cpu_time(lambda: od.apg_objective(P, prior, xc, K, FLAGS, sweeps))
        row.update({"cpu_B": cpu_B, "cpu_s": tc, "cpu_psw_s": S * cpu_B * sweeps / tc, "cpu_threads": torch.get_num_threads()})
    return row


def bmnist(S, B, sweeps, cpu_B, no_cpu, T=10):
    from apgs_b200.bmnist import apg_objective, init_models, Affine_Transformer
    from apgs_b200.resampler import Resampler
    from oracle import bmnist as ob
    K, FP, DP = 3, 96, 28
    torch.manual_seed(0)
    models = init_models(FP, DP, 400, 400, 2, 10, True, DEV)
    frames = ob.simulate(B, T, K)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, DP), torch.linspace(-1, 1, DP), indexing='ij')
    tmpl = torch.exp(-(xx ** 2 + yy ** 2) * 3).clamp(0, 1)
    AT = Affine_Transformer(FP, DP, True, DEV)
    rs = Resampler('systematic', S, True, DEV)
    fd = frames.to(DEV).unsqueeze(0).expand(S, -1, -1, -1, -1)
    mm = tmpl.to(DEV).repeat(S, B, K, 1, 1)
    with torch.no_grad():
        t = gpu_time(lambda: apg_objective(models, AT, fd, K, FLAGS, sweeps, rs, mm), iters=3, warm=1)
    row = {"task": "bmnist", "S": S, "B": B, "T": T, "sweeps": sweeps, "gpu_s": t, "gpu_psw_s": S * B * sweeps / t}
    if not no_cpu:
        torch.manual_seed(0)
        P, prior = ob.make_params(), ob.make_prior()
        fc = frames[:cpu_B].repeat(S, 1, 1, 1, 1)
        mc = tmpl.repeat(S, cpu_B, K, 1, 1)
        with torch.no_grad():
            tc = cpu_time(lambda: ob.apg_objective(P, prior, fc, K, FLAGS, sweeps, mc))
        row.update({"cpu_B": cpu_B, "cpu_s": tc, "cpu_psw_s": S * cpu_B * sweeps / tc, "cpu_threads": torch.get_num_threads()})
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--tasks", default="gmm,dmm,bmnist", help="comma-separated subset (the ncu captures of the C3 / C4 kernels)")
    a = ap.parse_args()
    tasks = set(a.tasks.split(","))
    torch.cuda.set_device(DEV)
    if WORLD > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=DEV)
    rows = []
    for fn, args in ((gmm, (10, 20, 5, 20)), (gmm, (1000, 1024, 10, 16)), (dmm, (100, 256, 7, 8)), (bmnist, (10, 64, 5, 2))):
        if fn.__name__ not in tasks:
            continue
        torch.manual_seed(1000 + RANK)                       # every rank owns different instances
        r = fn(*args, a.no_cpu or WORLD > 1 or RANK != 0)
        if WORLD > 1:                                        # slowest rank, N times the work (weak scaling along B)
            for key in ("gpu_s", "graph_s"):
                if key in r:
                    t = torch.tensor([r[key]], dtype=torch.float64, device=DEV)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    r[key] = float(t)
                    r[key.replace("_s", "_psw_s")] = WORLD * r["S"] * r["B"] * r["sweeps"] / r[key]
            r["n_gpus"], r["B_per_gpu"] = WORLD, r["B"]
        rows.append(r)
        if "cpu_psw_s" in r:
            r["gpu_over_cpu"] = r["gpu_psw_s"] / r["cpu_psw_s"]
        if RANK == 0:
            print(json.dumps(r), flush=True)
    if a.out and RANK == 0:
        json.dump(rows, open(a.out, "w"), indent=1)
    if WORLD > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

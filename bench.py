#!/usr/bin/env python
"""Benchmark of the APG sweep hot path (BASELINE.json metric: particle-sweeps/sec at GMM S=1000).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one apg_objective over one batch: GMM K=3 N=60 D=2, S=1000 particles x B=1024 instances
PER GPU (weak scaling along B, no data-path collective), 10 sweeps, block='decomposed', systematic
resampling, the reference's budget_analysis flags (ess + density, no loss).  value = particle-sweeps
per second = n_gpus * S * B * sweeps * K / time, time = max over ranks of the CUDA-event time of the
K steps.  Synthetic data drawn as apgs/gmm/sim_gmm.py; random-init encoder weights (default nn.Linear).

--impl reference: the reference's CPU path.  The reference is a Python package that does not exist
on the GPU box, so this arm runs the oracle port (oracle/gmm.py, verified bit-identical to the reference
under shared seeds in tests/test_oracle_golden.py, with mirror_cost=True so that every encoder is
evaluated twice exactly as the reference does) on all host cores, on a bounded sample of the same
workload (S=1000, B=--ref-batch, same sweeps).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

K_CL, D_X, N_PTS, H_Z = 3, 2, 60, 32
FLAGS = {'loss_required': False, 'ess_required': True, 'mode_required': False, 'density_required': True}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------- reference (CPU) arm
def cpu_port_throughput(S, B, sweeps, steps, warmup, seed=0):
    """particle-sweeps/s of the oracle port on the host cores (bounded sample)."""
    from oracle import gmm as og
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(seed)
    P, prior = og.make_params(K_CL, D_X, H_Z), og.make_prior(K_CL, D_X)
    x = og.simulate(B, N_PTS, K_CL, D_X).repeat(S, 1, 1, 1)
    with torch.no_grad():
        for _ in range(warmup):
            og.apg_objective(P, prior, x, FLAGS, sweeps, 'decomposed', 'systematic', mirror_cost=True)
        t0 = time.perf_counter()
        for _ in range(steps):
            og.apg_objective(P, prior, x, FLAGS, sweeps, 'decomposed', 'systematic', mirror_cost=True)
        dt = time.perf_counter() - t0
    return S * B * sweeps * steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S, B, sweeps = args.particles, args.ref_batch, args.sweeps
    v, sec = cpu_port_throughput(S, B, sweeps, args.steps, args.warmup)
    cores = torch.get_num_threads()
    sample = "GMM S=%d B=%d %d sweeps per step (of the B=%d workload), torch-CPU fp32, %d threads" % (S, B, sweeps, args.batch, cores)
    line = {"impl": "reference", "metric": "particle-sweeps/sec", "value": v, "unit": "particle-sweeps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": v, "unit": "particle-sweeps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "particle-sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": "GMM K=3 N=60 D=2, S=%d particles x B=%d instances per GPU, %d sweeps, decomposed, systematic"
                        % (args.particles, args.batch, args.sweeps),
            "S": args.particles, "B_per_gpu": args.batch, "num_sweeps": args.sweeps, "block": "decomposed",
            "resample_strategy": "systematic", "flags": "ess+density (budget_analysis)",
            "encoder": args.encoder, "parallelism": "instances sharded over ranks (dp%d), metric all-reduce only" % args.gpus,
            "cache": "inputs larger than L2 (z latent 737 MB per rank >> 126 MB L2); no flush"}


# ---------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist
    from apgs_b200 import ops, dist as adist
    from apgs_b200.gmm import apg_objective, init_apg_models
    from apgs_b200.resampler import Resampler
    from apgs_b200.gmm.sim import simulate

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    S, B, sweeps = args.particles, args.batch, args.sweeps

    torch.manual_seed(0)
    models = init_apg_models(K_CL, D_X, H_Z, True, dev)       # replicated weights (same seed on every rank)
    if args.encoder == "torch":
        models[1].fused = False
    resampler = Resampler('systematic', S, True, dev)
    torch.manual_seed(1 + rank)
    x_host = simulate(B, N_PTS, K_CL, D_X).pin_memory()    # this rank's shard of the instance batch
    x_dev = x_host.to(dev)

    def objective(x):
        return apg_objective(models, x, FLAGS, sweeps, 'decomposed', resampler)

    def step_eager(x):
        return adist.allreduce_metrics(objective(x))          # the only collective: (sweeps,3) all-reduce
    step = step_eager

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    graphed, launches_per_step = False, None
    with torch.no_grad():
        for _ in range(args.warmup):
            step(x_dev)
        if args.graph != "off":
            # the whole objective (all sweeps, ~800 launches) captured once into a CUDA graph and replayed per step;
            # random draws stay fresh (graph-safe Philox).  Falls back to eager launches if capture is refused.
            try:
                from apgs_b200.graph import GraphedObjective
                ops.PROFILE.reset(events=False)
                gobj = GraphedObjective(objective, x_dev, warmup=1)
                launches_per_step = ops.PROFILE.launches // 2          # 1 warm-up + 1 capture pass
                x_static = gobj.static_in[0]

                def step(x):                                            # noqa: F811
                    return adist.allreduce_metrics(gobj(x))
                for _ in range(args.warmup):
                    step(x_dev)
                graphed = True
            except Exception as exc:                                    # noqa: BLE001
                sys.stderr.write("bench: CUDA-graph capture failed (%r); running eager\n" % (exc,))
                step = step_eager
        # ---- device-resident throughput (value)
        sampler = ClockSampler(local)
        sampler.start()
        ops.PROFILE.reset(events=False)
        ms = timed(lambda: step(x_dev), args.steps)
        launches = launches_per_step * args.steps if graphed else ops.PROFILE.launches
        sampler.stop_flag = True
        # ---- end to end through the public API with HOST buffers (e2e)
        h2d = x_host.numel() * x_host.element_size()
        res_host = torch.empty(2, sweeps, dtype=torch.float64).pin_memory()

        def e2e_step():
            m = step(x_host.to(dev, non_blocking=True))
            res_host[0].copy_(m['ess'], non_blocking=True)
            res_host[1].copy_(m['density'], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        e2e_step()
        ms_e2e = timed(e2e_step, args.steps)
        # ---- per-kernel accounting for the roofline (separate pass: events around every C-ABI call)
        ops.PROFILE.reset(events=True)
        step_eager(x_dev)
        step_eager(x_dev)
        torch.cuda.synchronize()
        prof = ops.PROFILE.summary()
        dfr = ops.PROFILE.notes.get("distinct_ancestor_frac", [])
        distinct_frac = float(torch.stack(dfr).mean()) if dfr else None
        ops.PROFILE.reset(events=False)
        # ---- the same resample step on near-uniform weights (all ancestors distinct: every source row is read from HBM)
        probe = resample_probe(S, B, dev)
        final = step(x_dev)

    work = world * S * B * sweeps * args.steps
    value = work / (ms / 1e3)
    e2e = work / (ms_e2e / 1e3)
    peak, peak_src = peaks()
    g = prof.get("gather_ancestors", None)
    roof = None
    if g:
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_bench_traffic.json")))["k_gather"]["dram_bytes_per_launch"]
        except Exception:
            pass
        inner = (g["bytes_per_call"] / (S * B) - 4) / 2          # bytes of latent rows per particle actually gathered
        uniq = None if distinct_frac is None else S * B * (4 + inner * (1 + distinct_frac))
        roof = {"kernel": "k_gather (ancestor gather of tau, mu, z; the dominant kernel of the resample step)", "bound": "hbm",
                "achieved": g["gbs"], "peak": peak, "unit": "GB/s", "frac": g["gbs"] / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": g["bytes_per_call"],
                "avg_launch_ms": g["ms_avg"], "launches_timed": g["calls"],
                "distinct_ancestor_frac": distinct_frac,
                "achieved_unique_bytes": None if uniq is None else uniq / g["ms_avg"] / 1e6,
                "all_distinct_probe": probe,
                "latent_row_bytes_per_particle": inner,
                "note": "achieved = (4 + 2 * latent row bytes) per particle (index read + every row read and written once, at the "
                        "dtypes the step really moves) / CUDA-event time of each launch on the launching stream.  Inside the sweeps the "
                        "objective carries the assignments as uint8 labels, so a resampling step gathers 24 + 24 + 60 B per particle "
                        "instead of the reference layout's 24 + 24 + 720 B (SURVEY 8d: 1540 B): 225 MB per launch, which is launch- / "
                        "latency-bound (~70 us), and with random-init encoders the weights are degenerate (ESS ~ 1) so most reads hit "
                        "L2.  all_distinct_probe times the SAME kernels on the reference layout (1540 B per particle) with near-uniform "
                        "weights, where every row comes from HBM: that is the kernel's HBM roofline point."}
    zf = prof.get("gmm_z_step_fused", None)
    if roof is not None and zf:
        # the kernel with the largest share of the step is compute-bound: report it against ITS bound as well.  Floor =
        # tanh terms (S*B * 64 padded points * K * H) * 3.5 FMA-pipe cycles (14 packed instructions
        # of 2 cycles per 4 terms of 2 points) / (32 lanes * 4 schedulers * SMs) / SM clock sampled during the run.
        N_, K_, H_ = 60, 3, 32
        clk = (sampler.summary() or {}).get("sm_mhz") or 1965
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        terms = S * B * ((N_ + 31) // 32 * 32) * K_ * H_
        floor_ms = terms * 3.5 / 32 / (4 * sms) / (clk * 1e6) * 1e3
        tot = max(1e-9, sum(q["calls"] * q["ms_avg"] for q in prof.values()))
        roof["step_dominant_kernel"] = {
            "kernel": "k_gmm_z_fused (fused Enc_apg_z + draw + z-step weight)", "bound": "fp32 FMA pipe (not hbm, not tensor)",
            "share_of_step": round(zf["calls"] * zf["ms_avg"] / tot, 3), "avg_launch_ms": round(zf["ms_avg"], 4),
            "fma_pipe_floor_ms": round(floor_ms, 4), "frac_of_fma_pipe_floor": round(floor_ms / zf["ms_avg"], 3),
            "alg_GBs": round(zf["gbs"], 1), "frac_hbm": round(zf["gbs"] / peak, 3)}
    if rank == 0:
        line = {"metric": "particle-sweeps/sec", "value": value, "unit": "particle-sweeps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(workload_config(args), cuda_graph=graphed), "clocks": sampler.summary(),
                "e2e": {"value": e2e, "unit": "particle-sweeps/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": res_host.numel() * 8, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "roofline": roof,
                # every C-ABI call of one profiled step: share of the step's kernel time, algorithmic GB/s and its fraction
                # of the HBM peak.  gmm_z_step_fused (the largest share) is NOT an HBM-bound kernel: its inner loop sits at
                # the FP32 FMA-pipe bound, so its hbm fraction is low by construction; the roofline
                # object above is the gather, the HBM-bound kernel of the resample step the metric is about.
                "kernels": {k: {"calls": v["calls"], "ms_avg": round(v["ms_avg"], 4), "alg_GBs": round(v["gbs"], 1),
                                "frac_hbm": round(v["gbs"] / peak, 3),
                                "share": round(v["calls"] * v["ms_avg"] / max(1e-9, sum(q["calls"] * q["ms_avg"] for q in prof.values())), 3)}
                            for k, v in sorted(prof.items())},
                "final_metrics": {"ess_last": float(final['ess'][-1]), "density_last": float(final['density'][-1])}}
        if world == 1 and not args.no_cpu_baseline:
            v, sec = cpu_port_throughput(S, args.ref_batch, args.cpu_sweeps, 1, 0)
            line["cpu_baseline"] = {"value": v, "unit": "particle-sweeps/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "oracle port (bit-identical to the reference on CPU), GMM S=%d B=%d %d sweeps, "
                                              "one pass, %.1f s" % (S, args.ref_batch, args.cpu_sweeps, sec)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def resample_probe(S, B, dev, iters=10):
    """resample step (ancestors + gather of tau, mu, z) on log-weights ~ N(0, 0.1^2): ESS ~ S, so the gather reads
    every source row once -- algorithmic bytes == DRAM traffic."""
    from apgs_b200 import ops
    from apgs_b200.resampler import Resampler
    peak, _ = peaks()
    torch.manual_seed(123)
    lw = torch.randn(S, B, device=dev) * 0.1
    lat = [torch.randn(S, B, K_CL, D_X, device=dev), torch.randn(S, B, K_CL, D_X, device=dev),
           torch.randn(S, B, N_PTS, K_CL, device=dev)]
    r = Resampler('systematic', S, True, dev)
    u = r.draw_uniforms(S, B, dev)
    anc = ops.sample_ancestors(log_w=lw, strategy='systematic', uniforms=u, index_dtype=torch.int32)

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def med(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.zero_()        # evicts L2 and keeps the GPU busy while the host enqueues (no launch gap inside the events)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]
    t_g = med(lambda: ops.gather_ancestors(lat, anc))
    t_r = med(lambda: ops.resample(lat, log_w=lw, strategy='systematic', uniforms=u))
    nbytes = S * B * (4 + 2 * (2 * K_CL * D_X * 4 + N_PTS * K_CL * 4))
    return {"gather_ms": t_g, "gather_GBs": nbytes / t_g / 1e6, "gather_frac": nbytes / t_g / 1e6 / peak,
            "resample_step_ms": t_r, "resample_step_GBs": nbytes / t_r / 1e6, "resample_step_frac": nbytes / t_r / 1e6 / peak}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=1024, help="instances per GPU")
    ap.add_argument("--sweeps", type=int, default=10)
    ap.add_argument("--ref-batch", type=int, default=16, help="instances per step of the CPU arm")
    ap.add_argument("--cpu-sweeps", type=int, default=10)
    ap.add_argument("--encoder", default="auto", choices=["auto", "torch"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", default="auto", choices=["auto", "off"], help="replay the objective as one CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

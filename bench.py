#!/usr/bin/env python
"""Benchmark of the APG sweep hot path (BASELINE.json metric: particle-sweeps/sec at GMM S=1000).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-extras]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one apg_objective over one batch: GMM K=3 N=60 D=2, S=1000 particles x B=1024 instances
PER GPU (weak scaling along B, no data-path collective), 10 sweeps, block='decomposed', systematic
resampling, the reference's budget_analysis flags (ess + density, no loss).  value = particle-sweeps
per second = n_gpus * S * B * sweeps * K / time, time = max over ranks of the CUDA-event time of the
K steps.  Synthetic data drawn as apgs/gmm/sim_gmm.py; random-init encoder weights (default nn.Linear).

Besides the headline the line carries
  roofline      the step-dominant kernel (k_gmm_z_fused) against ITS bound -- the fp32 FMA pipe, it is neither HBM- nor
                tensor-bound -- and, under "hbm", the HBM-bound kernels of the resample step (in-step gather, the
                all-distinct probe on the reference layout, the eta weight kernel) against the measured copy peak;
  configs       the other BASELINE.json configs, driver-observed: C3 DMM (S=100, B=256, 7 sweeps), C4 bmnist (S=10, B=64,
                5 sweeps), C5 resampler microbench points, and (N > 1) a strong-scaling row of C2 (B=1024 in total);
                every entry with its dominant kernel, that kernel's roofline fraction and a CPU sample of the reference;
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, installed by __graft_entry__.build()) on the host cores.

--impl reference: the reference's own CPU implementation of the same path (apgs/gmm/objectives.py:9-53 through
oracle/ref_arm.py; `kind: "reference"`), all host threads, each step a bounded sample of the workload
(S=1000, B=--ref-batch instances, same sweeps); falls back to the oracle port (`kind: "port"`) only if oracle/_ref is
absent.
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
UNIT = "particle-sweeps/s"


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


def workload_config(args):
    return {"workload": "GMM K=3 N=60 D=2, S=%d particles x B=%d instances per GPU, %d sweeps, decomposed, systematic"
                        % (args.particles, args.batch, args.sweeps),
            "S": args.particles, "B_per_gpu": args.batch, "num_sweeps": args.sweeps, "block": "decomposed",
            "resample_strategy": "systematic", "flags": "ess+density (budget_analysis)",
            "encoder": args.encoder, "parallelism": "instances sharded over ranks (dp%d), metric all-reduce only" % args.gpus,
            "cache": "inputs larger than L2 (per-rank latents + per-step scratch ~0.4 GB >> 126 MB L2); no flush",
            "resampling": "19 ancestor draws per objective as in the reference; inside the sweeps the block updates read their "
                          "inputs through the (composed) ancestor maps instead of gathered copies -- bit-identical traces "
                          "(tests/test_gmm_lean_gpu.py); the gather kernel itself is measured by roofline.hbm.all_distinct_probe "
                          "and configs.C5_microbench",
            "ref_batch": args.ref_batch}


# ------------------------------------------------------------------------- reference (CPU) arm
def cpu_reference(task, S, B, sweeps, steps, warmup, threads=None):
    """(particle-sweeps/s, s/step, kind) of the reference's apg_objective for `task` on the host cores: the unmodified
    reference from oracle/_ref when it is installed, else the oracle port."""
    from oracle import ref_arm
    if ref_arm.available():
        fn = {"gmm": ref_arm.gmm_step_fn, "dmm": ref_arm.dmm_step_fn, "bmnist": ref_arm.bmnist_step_fn}[task](S, B, sweeps)
        kind = "reference"
    else:
        fn, kind = _port_step_fn(task, S, B, sweeps), "port"
    v, sec = ref_arm.throughput(fn, S * B * sweeps, steps, warmup, threads)
    return v, sec, kind


def _port_step_fn(task, S, B, sweeps):
    torch.manual_seed(0)
    if task == "gmm":
        from oracle import gmm as og
        P, prior = og.make_params(K_CL, D_X, H_Z), og.make_prior(K_CL, D_X)
        x = og.simulate(B, N_PTS, K_CL, D_X).repeat(S, 1, 1, 1)
        return lambda: og.apg_objective(P, prior, x, FLAGS, sweeps, 'decomposed', 'systematic', mirror_cost=True)
    if task == "dmm":
        from oracle import dmm as od
        P, prior = od.make_params(4, 2), od.make_prior(4, 2)
        x = od.simulate(B, 200, 4, 2).repeat(S, 1, 1, 1)
        return lambda: od.apg_objective(P, prior, x, 4, FLAGS, sweeps)
    from oracle import bmnist as ob
    P, prior = ob.make_params(), ob.make_prior()
    fr = ob.simulate(B, 10, 3).repeat(S, 1, 1, 1, 1)
    mm = _blob().repeat(S, B, 3, 1, 1)
    return lambda: ob.apg_objective(P, prior, fr, 3, FLAGS, sweeps, mm)


def _blob(DP=28):
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, DP), torch.linspace(-1, 1, DP), indexing='ij')
    return torch.exp(-(xx ** 2 + yy ** 2) * 3).clamp(0, 1)


def cpu_extras(args):
    """CPU samples of the other configs (reference on the host cores, bounded)."""
    out = {}
    with torch.no_grad():
        v, sec, kind = cpu_reference("dmm", 100, 8, 7, 1, 1)
        out["C3_dmm"] = {"value": v, "unit": UNIT, "kind": kind, "cores": torch.get_num_threads(),
                         "sample": "DMM K=4 N=200 S=100 B=8 7 sweeps, one pass, %.1f s" % sec}
        v, sec, kind = cpu_reference("bmnist", 10, 2, 5, 1, 1)
        out["C4_bmnist"] = {"value": v, "unit": UNIT, "kind": kind, "cores": torch.get_num_threads(),
                            "sample": "bmnist K=3 T=10 96x96 S=10 B=2 5 sweeps, one pass, %.1f s" % sec}
    from oracle import ref_arm
    if ref_arm.available():
        rows = []
        for strat in ("systematic", "multinomial"):
            for d in (2, 32, 256):
                fn = ref_arm.resample_step_fn(1000, 1000, d, strat)
                v, sec = ref_arm.throughput(fn, 1000 * 1000, 3, 1)
                rows.append({"strategy": strat, "P": 10**6, "S": 1000, "d": d, "particles_per_s": v,
                             "alg_GBs": v * (4 + 8 * d) / 1e9, "ms": sec * 1e3})
        out["C5_microbench"] = {"kind": "reference", "cores": torch.get_num_threads(), "points": rows,
                                "sample": "Resampler.sample_ancestral_index + resample_4dims on CPU, S=1000 B=1000, 3 passes"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S, B, sweeps = args.particles, args.ref_batch, args.sweeps
    with torch.no_grad():
        # warm-up passes on a smaller batch (they only warm the thread pool, the allocator and the caches; CPU throughput is
        # flat in B), timed steps on the full sample: keeps the whole arm within a few minutes at --steps 20 --warmup 5
        if args.warmup > 0:
            cpu_reference("gmm", S, max(1, min(B, 8)), sweeps, max(1, args.warmup - 1), 0)
        # one calibration step at B = 8 sizes the timed sample so that the --steps run stays within --ref-seconds of CPU work
        # (the contract's "bounded sample"): B = --ref-batch (64) when it fits, otherwise the largest multiple of 8
        _, sec8, _ = cpu_reference("gmm", S, max(1, min(B, 8)), sweeps, 1, 0)
        fit = int(args.ref_seconds / max(1e-6, args.steps * sec8 / max(1, min(B, 8))))
        B = max(8, min(B, fit // 8 * 8)) if B > 8 else B
        v, sec, kind = cpu_reference("gmm", S, B, sweeps, args.steps, 0)
        cores = torch.get_num_threads()
        v1, sec1, _ = cpu_reference("gmm", S, 4, sweeps, 1, 0, threads=1)
    torch.set_num_threads(cores)
    sample = "GMM S=%d B=%d %d sweeps per timed step (a bounded sample of the B=%d workload; CPU throughput is flat in B; " \
             "warm-up steps at B=%d), torch-CPU fp32, %d threads" % (S, B, sweeps, args.batch, max(1, min(B, 8)), cores)
    base = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
            "value_1thread": v1, "sample_1thread": "S=%d B=4 %d sweeps, one pass, %.1f s" % (S, sweeps, sec1)}
    line = {"impl": "reference", "metric": "particle-sweeps/sec", "value": v, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args), ref_batch=B), "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_extras:
        line["configs"] = cpu_extras(args)
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------- our arm
class Ctx(object):
    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, k):
        """ms of k calls of fn: barrier + synchronize on both sides, CUDA events on the launching stream, max over ranks."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())


def kernel_table(prof, peak):
    tot = max(1e-9, sum(q["calls"] * q["ms_avg"] for q in prof.values()))
    return {k: {"calls": v["calls"], "ms_avg": round(v["ms_avg"], 4), "alg_GBs": round(v["gbs"], 1),
                "frac_hbm": round(v["gbs"] / peak, 3), "share": round(v["calls"] * v["ms_avg"] / tot, 3)}
            for k, v in sorted(prof.items())}


def profile_step(fn, passes=2):
    """per-kernel accounting: CUDA events around every C-ABI call of `passes` eager steps."""
    from apgs_b200 import ops
    ops.PROFILE.reset(events=True)
    for _ in range(passes):
        fn()
    torch.cuda.synchronize()
    prof = ops.PROFILE.summary()
    notes = ops.PROFILE.notes
    ops.PROFILE.reset(events=False)
    return prof, notes


def run_ours(args):
    from apgs_b200 import ops, dist as adist
    from apgs_b200.gmm import apg_objective, apg_objective_streams, init_apg_models
    from apgs_b200.resampler import Resampler
    from apgs_b200.gmm.sim import simulate

    cx = Ctx()
    world, rank, dev = cx.world, cx.rank, cx.dev
    S, B, sweeps = args.particles, args.batch, args.sweeps
    peak, peak_src = peaks()

    torch.manual_seed(0)
    models = init_apg_models(K_CL, D_X, H_Z, True, dev)       # replicated weights (same seed on every rank)
    if args.encoder == "torch":
        models[1].fused = False
    resampler = Resampler('systematic', S, True, dev)
    torch.manual_seed(1 + rank)
    x_host = simulate(B, N_PTS, K_CL, D_X).pin_memory()    # this rank's shard of the instance batch
    x_dev = x_host.to(dev)

    def objective(x):
        if args.streams > 1:
            return apg_objective_streams(models, x, FLAGS, sweeps, 'decomposed', resampler, parts=args.streams)
        return apg_objective(models, x, FLAGS, sweeps, 'decomposed', resampler)

    def step_eager(x):
        return adist.allreduce_metrics(objective(x))          # the only collective: (sweeps,4) all-reduce
    step = step_eager

    graphed, launches_per_step = False, None
    with torch.no_grad():
        for _ in range(args.warmup):
            step(x_dev)
        if args.graph != "off":
            # the whole objective (all sweeps) captured once into a CUDA graph and replayed per step;
            # random draws stay fresh (graph-safe Philox).  Falls back to eager launches if capture is refused.
            try:
                from apgs_b200.graph import GraphedObjective
                ops.PROFILE.reset(events=False)
                gobj = GraphedObjective(objective, x_dev, warmup=1)
                launches_per_step = ops.PROFILE.launches // 2          # 1 warm-up + 1 capture pass

                def step(x):                                            # noqa: F811
                    return adist.allreduce_metrics(gobj(x))
                for _ in range(args.warmup):
                    step(x_dev)
                graphed = True
            except Exception as exc:                                    # noqa: BLE001
                sys.stderr.write("bench: CUDA-graph capture failed (%r); running eager\n" % (exc,))
                step = step_eager
        # ---- device-resident throughput (value)
        sampler = ClockSampler(cx.local)
        sampler.start()
        ops.PROFILE.reset(events=False)
        ms = cx.timed(lambda: step(x_dev), args.steps)
        launches = launches_per_step * args.steps if graphed else ops.PROFILE.launches
        sampler.stop_flag = True
        # ---- end to end through the public API with HOST buffers (e2e)
        h2d = x_host.numel() * x_host.element_size()
        res_host = torch.empty(3, sweeps, dtype=torch.float64).pin_memory()

        def e2e_step():
            m = step(x_host.to(dev, non_blocking=True))
            res_host[0].copy_(m['ess'], non_blocking=True)
            res_host[1].copy_(m['density'], non_blocking=True)
            res_host[2].copy_(m['log_z'], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        e2e_step()
        ms_e2e = cx.timed(e2e_step, args.steps)
        # ---- per-kernel accounting for the roofline (separate pass: events around every C-ABI call)
        prof, notes = profile_step(lambda: step_eager(x_dev))
        dfr = notes.get("distinct_ancestor_frac", [])
        distinct_frac = float(torch.stack(dfr).mean()) if dfr else None
        # ---- the same resample step on near-uniform weights (all ancestors distinct: every source row is read from HBM)
        probe = resample_probe(S, B, dev)
        final = step(x_dev)
        final = {k: v.clone() for k, v in final.items()}

        work = world * S * B * sweeps * args.steps
        value = work / (ms / 1e3)
        e2e = work / (ms_e2e / 1e3)
        clocks = sampler.summary()
        roof = gmm_roofline(prof, S, B, peak, peak_src, clocks, distinct_frac, probe, dev)
        line = {"metric": "particle-sweeps/sec", "value": value, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args), "cuda_graph": graphed, "clocks": clocks,
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": res_host.numel() * 8, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "roofline": roof,
                # every C-ABI call of one profiled (eager) step: share of the step's kernel time, algorithmic GB/s and its
                # fraction of the HBM peak
                "kernels": kernel_table(prof, peak),
                "final_metrics": {"ess_last": float(final['ess'][-1]), "density_last": float(final['density'][-1]),
                                  "log_z_last": float(final['log_z'][-1])}}
        if not args.no_extras:
            line["configs"] = extras(cx, args, peak)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            with torch.no_grad():
                v, sec, kind = cpu_reference("gmm", S, args.ref_batch, args.cpu_sweeps, 1, 0)
                cores = torch.get_num_threads()
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                                        "sample": "GMM S=%d B=%d %d sweeps, one pass, %.1f s, torch-CPU fp32, %d threads"
                                                  % (S, args.ref_batch, args.cpu_sweeps, sec, cores)}
                if not args.no_extras:
                    for k, v in cpu_extras(args).items():
                        if k in line["configs"]:
                            line["configs"][k]["cpu_baseline"] = v
                    line["reference_cuda"] = reference_on_gpu(S, dev)
        print(json.dumps(line), flush=True)
    if world > 1:
        cx.dist.destroy_process_group()


def gmm_roofline(prof, S, B, peak, peak_src, clocks, distinct_frac, probe, dev):
    """roofline object: the kernel with the largest share of the step against its own bound, then the HBM-bound ones."""
    zf, g, ew = prof.get("gmm_z_step_fused"), prof.get("gather_ancestors"), prof.get("gmm_logw_eta_stats")
    if not zf:
        return None
    # k_gmm_z_fused: S*B particles x 64 padded point slots x K x H tanh terms; the minimal formulation this repo found
    # costs 3.5 FMA-pipe lane-cycles per term (14 packed FFMA2/FMUL2 of 2 cycles each per 4 hidden
    # units of 2 points).  achieved = those lane-cycles / launch time; peak = SMs x 128 FMA lanes x SM clock (sampled):
    # frac is the share of the kernel's time the FMA pipe would need for the inner loop alone.
    clk = (clocks or {}).get("sm_mhz") or 1965
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    terms = S * B * ((N_PTS + 31) // 32 * 32) * K_CL * H_Z
    fma_ops = terms * 3.5                                            # FFMA2 / FMUL2 occupy a lane for 2 cycles (2 ops)
    peak_fma = sms * 128 * clk * 1e6 / 1e12                          # T FMA-lane-cycles / s
    ach = fma_ops / (zf["ms_avg"] * 1e-3) / 1e12
    tot = max(1e-9, sum(q["calls"] * q["ms_avg"] for q in prof.values()))
    roof = {"kernel": "k_gmm_z_fused (fused Enc_apg_z MLP + softmax + draw + z-step weight): the step-dominant kernel",
            "bound": "fp32 FMA pipe / instruction issue (compute; neither hbm nor tensor -- nothing on the path is a dense "
                     "contraction; packed FFMA2 / FMUL2 take two issue slots, see mlp_loop_alone)",
            "achieved": ach, "peak": peak_fma, "unit": "T lane-cycles/s of the fp32 FMA pipe (inner loop: 3.5 per tanh term)",
            "frac": ach / peak_fma, "traffic": None,
            "traffic_note": "DRAM bytes are not measurable inside the run; see ncu_capture (the committed ncu --set full capture "
                            "of this command: k_gmm_z_fused moves ~0.2 GB per launch, FMA pipe cycles active 68 %, issue active 68 % with packed "
                            "FFMA2 / FMUL2 counted once)",
            "share_of_step_kernel_time": round(zf["calls"] * zf["ms_avg"] / tot, 3), "avg_launch_ms": zf["ms_avg"],
            "launches_timed": zf["calls"], "sm_mhz": clk, "alg_GBs": zf["gbs"], "frac_hbm": zf["gbs"] / peak,
            "peak_source": "SMs x 128 lanes x sampled SM clock; hbm: " + peak_src}
    roof["mlp_loop_alone"] = mlp_loop_probe(S, B, dev, zf["ms_avg"], fma_ops / (peak_fma * 1e12) * 1e3)
    cap = ncu_capture()
    if cap:
        roof["ncu_capture"] = cap
    hbm = {"peak": peak, "unit": "GB/s", "peak_source": peak_src}
    if g:
        inner = (g["bytes_per_call"] / (S * B) - 4) / 2          # bytes of latent rows per particle actually gathered
        hbm["gather_in_step"] = {
            "kernel": "k_gather on the label-mode latents inside the sweep", "avg_launch_ms": g["ms_avg"],
            "algorithmic_bytes_per_launch": g["bytes_per_call"], "alg_GBs": g["gbs"], "frac_algorithmic": g["gbs"] / peak,
            "latent_row_bytes_per_particle": inner, "distinct_ancestor_frac": distinct_frac,
            "unique_bytes_GBs": None if distinct_frac is None else S * B * (4 + inner * (1 + distinct_frac)) / g["ms_avg"] / 1e6,
            "note": "random-init encoders give degenerate weights (ESS ~ 1): almost every source row is an L2 hit, so the "
                    "algorithmic rate is not a DRAM rate (ncu dram bytes ~0.26x of algorithmic, profiles/); the HBM point "
                    "of this kernel is all_distinct_probe"}
    hbm["all_distinct_probe"] = probe
    if ew:
        hbm["eta_weights"] = {"kernel": "k_gmm_eta_weights_stats (+ fused draw)", "avg_launch_ms": ew["ms_avg"],
                              "alg_GBs": ew["gbs"], "frac": ew["gbs"] / peak}
    roof["hbm"] = hbm
    return roof


def mlp_loop_probe(S, B, dev, step_ms, fma_floor_ms, iters=5):
    """The z step's main loop ALONE (apgs_gmm_z_loop_probe_f32: the kernel's own mlp_quad_fast on synthetic shared-memory
    tables, no prologue / epilogue), timed live for the loop work of one launch (S*B/32 CTAs x 2 passes): the ceiling of
    the MLP part on this GPU.  Packed FFMA2 / FMUL2 take two issue slots, so the loop (84 packed + 27 other instructions per
    quad of hidden units) is bound by instruction issue above its FMA-pipe time; the step's other ~1500 instructions per
    warp share the same issue port."""
    from apgs_b200.gmm import kernels
    ctas = (S * B + 31) // 32
    out = torch.empty(ctas * 256, dtype=torch.float32, device=dev)
    kernels.z_loop_probe(ctas, 2, out)
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        kernels.z_loop_probe(ctas, 2, out)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    return {"ms": ms, "fma_pipe_floor_ms": fma_floor_ms, "loop_ceiling_frac_of_fma_pipe": fma_floor_ms / ms,
            "step_launch_ms": step_ms, "loop_share_of_step_launch": ms / step_ms,
            "note": "the loop alone, same work as one launch of the step, measured in this run; FFMA2/FMUL2 occupy two issue "
                    "slots each (benchmarks/micro/mlp_loop.cu, profiles/r2_z_step_issue_bound.md)"}


def ncu_capture():
    """DRAM bytes per launch of the kernels of this command from the COMMITTED `ncu --set full` capture (profiles/
    ncu_bench_traffic.json, written by benchmarks/summarize_profiles.py) -- reported beside the live numbers, never as
    them: `roofline.traffic` stays null because DRAM bytes cannot be measured inside the run."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_bench_traffic.json")))
    except Exception:
        return None
    out = {"source": "profiles/ncu_bench_traffic.json + profiles/r2_ncu_kernels.md (committed capture, not measured in this run)"}
    for k in ("k_gmm_z_fused", "k_gather", "probe:k_gather", "k_gmm_nss_im", "k_gmm_eta_weights_stats"):
        if k in t:
            out[k] = {"dram_bytes_per_launch": t[k]["dram_bytes_per_launch"], "duration_us": t[k]["duration_us"],
                      "dram_GBs": t[k]["dram_bytes_per_launch"] / t[k]["duration_us"] / 1e3}
    return out


def resample_probe(S, B, dev, iters=10):
    """resample step (ancestors + gather of tau, mu, z) on log-weights ~ N(0, 0.1^2), reference layout (1540 B per
    particle): ESS ~ S, so the gather reads every source row once -- algorithmic bytes == DRAM traffic; L2 flushed
    between iterations."""
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
    t_g = _median_ms(lambda: ops.gather_ancestors(lat, anc), flush, iters)
    t_r = _median_ms(lambda: ops.resample(lat, log_w=lw, strategy='systematic', uniforms=u), flush, iters)
    nbytes = S * B * (4 + 2 * (2 * K_CL * D_X * 4 + N_PTS * K_CL * 4))
    return {"layout": "reference layout: tau, mu (24 B) + one-hot z (720 B) per particle, near-uniform weights, L2 flushed",
            "algorithmic_bytes": nbytes, "gather_ms": t_g, "gather_GBs": nbytes / t_g / 1e6, "gather_frac": nbytes / t_g / 1e6 / peak,
            "resample_step_ms": t_r, "resample_step_GBs": nbytes / t_r / 1e6, "resample_step_frac": nbytes / t_r / 1e6 / peak}


def _median_ms(fn, flush, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if flush is not None:
            flush.zero_()    # evicts L2 and keeps the GPU busy while the host enqueues (no launch gap inside the events)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def reference_on_gpu(S, dev, B=64, sweeps=10):
    """the UNMODIFIED reference through stock PyTorch CUDA ops on the same GPU ("second, stronger
    baseline"), bounded sample."""
    from oracle import ref_arm
    if not ref_arm.available():
        return {"unavailable": "oracle/_ref not installed"}
    try:
        fn = ref_arm.gmm_step_fn(S, B, sweeps, device=str(dev))
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / 3
        return {"value": S * B * sweeps / sec, "unit": UNIT, "ms_per_step": sec * 1e3,
                "sample": "apgs.gmm.objectives.apg_objective on cuda, S=%d B=%d %d sweeps, 3 passes" % (S, B, sweeps)}
    except Exception as exc:                                            # noqa: BLE001
        return {"unavailable": repr(exc)[:200]}


# ------------------------------------------------------------------------------------- the other configs
def _dominant(prof, peak):
    tab = kernel_table(prof, peak)
    top = sorted(tab.items(), key=lambda kv: -kv[1]["share"])[:3]
    return [dict(kernel=k, **v) for k, v in top]


def _graphed(objective, x, eager):
    """the objective captured once into a CUDA graph and replayed (as the C2 step is); eager launches if capture is refused"""
    try:
        from apgs_b200.graph import GraphedObjective
        g = GraphedObjective(objective, x, warmup=1)
        xin = g.static_in[0]
        run = lambda: g(xin)                                        # noqa: E731
        run()
        torch.cuda.synchronize()
        return run, True
    except Exception as exc:                                        # noqa: BLE001
        sys.stderr.write("bench: CUDA-graph capture of an extra config failed (%r); running eager\n" % (exc,))
        torch.cuda.synchronize()
        return eager, False


def extras(cx, args, peak):
    from apgs_b200.resampler import Resampler
    dev, world, rank = cx.dev, cx.world, cx.rank
    out = {}
    # ---- C3: DMM ring clustering K=4, N=200, S=100, B=256 per GPU, 7 sweeps
    from apgs_b200.dmm import apg_objective as dmm_objective, init_apg_models as dmm_models
    from apgs_b200.dmm.sim import simulate as dmm_sim
    K, D, N, S, B, sweeps = 4, 2, 200, 100, 256, 7
    torch.manual_seed(0)
    models = dmm_models(K, D, 32, 8, 32, 32, 0.5, True, dev)
    torch.manual_seed(100 + rank)
    x = dmm_sim(B, N, K, D).to(dev)
    rs = Resampler('systematic', S, True, dev)
    fn = lambda: dmm_objective(models, x, K, FLAGS, sweeps, rs)        # noqa: E731
    for _ in range(3):
        fn()
    run, c3_graph = _graphed(lambda xx: dmm_objective(models, xx, K, FLAGS, sweeps, rs), x, fn)
    ms = cx.timed(run, 5) / 5
    prof, _ = profile_step(fn, 1)
    top = _dominant(prof, peak)
    roof = {"kernel": top[0]["kernel"], "bound": "hbm", "achieved": top[0]["alg_GBs"], "peak": peak, "unit": "GB/s", "frac": top[0]["frac_hbm"]}
    le = prof.get("dmm_local_encoder_fused")
    if le and top[0]["kernel"] == "dmm_local_encoder_fused":
        # k_dmm_local_fused is a fused MLP, not a stream: per point K*H ReLU units of the pi MLP (3 FMA-pipe lane-cycles: first
        # layer, max, output FMA) and 2 nets x 2 assignments x H tanh units of the Beta-concentration MLPs (4.25: pair inversion)
        H = 32
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        cyc = S * B * N * (K * H * 3.0 + 4 * H * 4.25)
        ach = cyc / (le["ms_avg"] * 1e-3) / 1e12
        pk = sms * 128 * 1965e6 / 1e12
        roof = {"kernel": "k_dmm_local_fused", "bound": "fp32 FMA pipe (fused per-point MLPs)", "achieved": ach, "peak": pk,
                "unit": "T lane-cycles/s", "frac": ach / pk, "alg_GBs": le["gbs"], "frac_hbm": le["gbs"] / peak}
    out["C3_dmm"] = {"workload": "DMM K=4 N=200 D=2, S=100 x B=256 per GPU, 7 sweeps, systematic", "value": world * S * B * sweeps / (ms / 1e3),
                     "unit": UNIT, "ms_per_step": ms, "n_gpus": world, "scaling": "weak", "cuda_graph": c3_graph,
                     "dominant_kernels": top, "roofline": roof}
    del models, x, run
    # ---- C4: bouncing MNIST K=3 digits, T=10 frames of 96x96, S=10, B=64 per GPU, 5 sweeps
    from apgs_b200.bmnist import apg_objective as bm_objective, init_models as bm_models, Affine_Transformer
    from apgs_b200.bmnist.sim import simulate as bm_sim
    K, T, S, B, sweeps, FP, DP = 3, 10, 10, 64, 5, 96, 28
    torch.manual_seed(0)
    models = bm_models(FP, DP, 400, 400, 2, 10, True, dev)
    torch.manual_seed(200 + rank)
    frames0 = bm_sim(B, T, K).to(dev)
    frames = frames0.unsqueeze(0).expand(S, -1, -1, -1, -1)            # stride-0 S axis, never materialised
    AT = Affine_Transformer(FP, DP, True, dev)
    mm = _blob(DP).to(dev).repeat(S, B, K, 1, 1)
    rs = Resampler('systematic', S, True, dev)
    fn = lambda: bm_objective(models, AT, frames, K, FLAGS, sweeps, rs, mm)     # noqa: E731
    for _ in range(2):
        fn()
    run, c4_graph = _graphed(lambda f0: bm_objective(models, AT, f0.unsqueeze(0).expand(S, -1, -1, -1, -1), K, FLAGS, sweeps, rs, mm),
                             frames0, fn)
    ms = cx.timed(run, 3) / 3
    prof, _ = profile_step(fn, 1)
    top = _dominant(prof, peak)
    tm = prof.get("bmnist_template_match")
    roof = {"kernel": top[0]["kernel"], "bound": "hbm", "achieved": top[0]["alg_GBs"], "peak": peak, "unit": "GB/s",
            "frac": top[0]["frac_hbm"]}
    if tm and top[0]["kernel"] == "bmnist_template_match":
        # S*B independent 96x96 (*) 28x28 correlations per call: 69^2 * 28^2 FMA each -> FMA-pipe work, not HBM
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        fma = S * B * (FP - DP + 1) ** 2 * DP * DP
        ach = fma / (tm["ms_avg"] * 1e-3) / 1e12
        pk = sms * 128 * 1965e6 / 1e12
        roof = {"kernel": "k_template_match", "bound": "fp32 FMA pipe", "achieved": ach, "peak": pk, "unit": "TFMA/s", "frac": ach / pk}
    out["C4_bmnist"] = {"workload": "bmnist K=3 T=10 96x96, S=10 x B=64 per GPU, 5 sweeps, systematic", "value": world * S * B * sweeps / (ms / 1e3),
                        "unit": UNIT, "ms_per_step": ms, "n_gpus": world, "scaling": "weak", "cuda_graph": c4_graph,
                        "dominant_kernels": top, "roofline": roof}
    del models, frames, frames0, mm, run
    torch.cuda.empty_cache()
    # ---- C5: resampler microbench, one latent (S,B,1,d) fp32 per GPU; algorithmic bytes = P (4 + 8 d)
    out["C5_microbench"] = microbench(cx, peak)
    # ---- strong scaling of C2: B = 1024 instances IN TOTAL, split over the ranks (per-GPU floor of the step)
    if world > 1:
        from apgs_b200 import dist as adist
        from apgs_b200.gmm import apg_objective, init_apg_models
        from apgs_b200.gmm.sim import simulate
        S, Bt, sweeps = args.particles, args.batch, args.sweeps
        lo, hi = adist.shard_bounds(Bt, rank, world)
        torch.manual_seed(0)
        models = init_apg_models(K_CL, D_X, H_Z, True, dev)
        x = simulate(Bt, N_PTS, K_CL, D_X)[lo:hi].to(dev)
        rs = Resampler('systematic', S, True, dev)
        fn = lambda: adist.allreduce_metrics(apg_objective(models, x, FLAGS, sweeps, 'decomposed', rs))   # noqa: E731
        for _ in range(3):
            fn()
        ms = cx.timed(fn, 5) / 5
        out["C2_strong"] = {"workload": "GMM C2 with B=%d instances in total (B/N per GPU), eager launches" % Bt,
                            "value": S * Bt * sweeps / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "n_gpus": world, "scaling": "strong"}
    return out


def microbench(cx, peak):
    from apgs_b200 import ops
    from apgs_b200.resampler import Resampler
    dev, world = cx.dev, cx.world
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    rows = []
    for P in (10**6, 10**8):
        S = 1000
        B = P // S
        for d in (2, 32, 256):
            if 2 * P * d * 4 > 100e9:
                rows.append({"P_per_gpu": P, "S": S, "d": d, "skipped": "2 x %.0f GB of latent rows exceed the HBM budget of one GPU" % (P * d * 4 / 1e9)})
                continue
            lw = torch.randn(S, B, device=dev) * 0.1     # ESS ~ S: (nearly) every source row is read, algorithmic == DRAM bytes
            lat = [torch.randn(S, B, 1, d, device=dev)]
            for strat in ("systematic", "multinomial"):
                r = Resampler(strat, S, True, dev)
                u = r.draw_uniforms(S, B, dev)
                lay = "bs" if strat == "multinomial" else None          # draw_uniforms: (B,S), the reference's generator order
                fn = lambda: ops.resample(lat, log_w=lw, strategy=strat, uniforms=u, uniform_layout=lay)      # noqa: E731
                cx.barrier()
                t = torch.tensor([_median_ms(fn, flush, iters=7, warm=2)], dtype=torch.float64, device=dev)
                if world > 1:
                    cx.dist.all_reduce(t, op=cx.dist.ReduceOp.MAX)
                t = float(t)
                nbytes = P * (4 + 8 * d)
                rows.append({"strategy": strat, "P_per_gpu": P, "S": S, "B_per_gpu": B, "d": d, "ms": round(t, 4),
                             "particles_per_s": world * P / (t / 1e3), "alg_GBs_per_gpu": round(nbytes / t / 1e6, 1),
                             "frac_hbm": round(nbytes / t / 1e6 / peak, 3)})
            del lw, lat
            torch.cuda.empty_cache()
    return {"workload": "resample (ancestors + gather) of one latent (S,B,1,d) fp32, log_w ~ N(0, 0.1^2) (ESS ~ S, every source row read), L2 flushed between "
                        "iterations, median of 7; weak scaling (P particles per GPU)", "n_gpus": world, "points": rows,
            "roofline": {"bound": "hbm", "peak": peak, "unit": "GB/s",
                         "frac_best": max([r.get("frac_hbm", 0) for r in rows] or [0])}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=1024, help="instances per GPU")
    ap.add_argument("--sweeps", type=int, default=10)
    ap.add_argument("--ref-batch", type=int, default=64, help="instances per step of the CPU arm (SURVEY.md 8d)")
    ap.add_argument("--ref-seconds", type=float, default=150.0,
                    help="--impl reference: budget of CPU seconds for the timed steps; the sample is shrunk below --ref-batch to fit")
    ap.add_argument("--cpu-sweeps", type=int, default=10)
    ap.add_argument("--encoder", default="auto", choices=["auto", "torch"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the C3 / C4 / C5 / strong-scaling entries of `configs`")
    ap.add_argument("--streams", type=int, default=1, help="slices of the instance batch run on separate CUDA streams")
    ap.add_argument("--graph", default="auto", choices=["auto", "off"], help="replay the objective as one CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
